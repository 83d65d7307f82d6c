#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest10.log; tail -3 gpurun_out/r2_pytest10.log
python -m pytest tests/test_gpu_sampler_extras.py -m gpu -q -s -k "long_trajectory" 2>&1 | grep "long trajectory" | tee gpurun_out/r2_long_trajectory.log
for v in 2 1; do
  echo "== VAVAE_ATTN_PER_SM=$v"
  VAVAE_ATTN_PER_SM=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep attention
done 2>&1 | tee gpurun_out/r2_ab_attn_persm.log
