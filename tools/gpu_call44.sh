#!/bin/bash
# B/1 mlp.w3 (K = 2048): register-prefetch epilogue (default) vs TMA epilogue on eight / four warps, now with the tile-ahead vectors
for cfg in "2 0" "1 8" "1 4"; do
  set -- $cfg
  echo "== TMA_EPI=$1 XT_WARPS=$2"
  VAVAE_GEMM_TMA_EPI=$1 VAVAE_GEMM_XT_WARPS=$2 timeout -s KILL 100 python tools/ab_gemms.py 2>&1 | grep "B/1 gemm_w3\|B/1 gemm_proj"
done
