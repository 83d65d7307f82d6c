#!/bin/bash
# What bounds the gated-residual GEMMs (proj is HBM-heavy)?  Timing-only experiments: drop the residual loads (1), the
# next-operand stores (2), the residual stores (4), and combinations.
for v in 0 1 2 4 3 7; do
  echo "== VAVAE_GEMM_EXPERIMENT=$v"
  VAVAE_GEMM_EXPERIMENT=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep -E "proj|w3"
done 2>&1 | tee gpurun_out/r2_ab_proj_experiment.log
timeout -s KILL 600 python tools/torch_gpu_baseline.py xl1 --compile 2>&1 | tail -8 | tee gpurun_out/r2_torch_baseline_xl1.log
timeout -s KILL 400 python tools/torch_gpu_baseline.py b1 --compile 2>&1 | tail -8 | tee gpurun_out/r2_torch_baseline_b1.log
