#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest3.log
tail -8 gpurun_out/r2_pytest3.log
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_ATTN_EXPERIMENT=$v (rep $rep)"
  VAVAE_ATTN_EXPERIMENT=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep attention
done; done 2>&1 | tee gpurun_out/r2_ab_attn_nomax.log
