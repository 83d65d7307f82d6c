#!/bin/bash
# Is the GEMM operand feed limited per SM (ingress port) or chip-wide (L2 / crossbar)?  Ingress-heavy 128-wide tiles on
# 148 vs 74 vs 36 SMs: a chip-wide limit keeps total throughput, a per-SM limit halves it.
for bn in 128 256; do for p in 0 37 18; do
  echo "== BN=$bn max_pairs=$p"
  VAVAE_GEMM_BN=$bn VAVAE_GEMM_MAX_PAIRS=$p AB_MODELS=xl timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep -E "w12|qkv"
done; done 2>&1 | tee gpurun_out/r2_ab_pairs.log
