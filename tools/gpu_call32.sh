#!/bin/bash
# narrow tiles for small-M launches: correctness (batch invariance, fixtures), then the batch sweep with and without
echo "== parity tests"
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_epilogues.py -m gpu -q -x 2>&1 | tail -5
for v in 0 1; do
  echo "== VAVAE_GEMM_NARROW=$v"
  VAVAE_GEMM_NARROW=$v timeout -s KILL 300 python tools/microbench_forward.py --batches 1,2,4,8,16,32 --iters 20 2>&1 | cut -c1-150
done
