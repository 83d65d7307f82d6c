#!/bin/bash
for rep in 1 2; do for v in 0 2; do
  echo "== VAVAE_ATTN_EXPERIMENT=$v (rep $rep)"
  VAVAE_ATTN_EXPERIMENT=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep attention
done; done 2>&1 | tee gpurun_out/r2_ab_attn_spin.log
python -m pytest tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -2
