#!/bin/bash
# how much of the TMA epilogue's loss on mlp.w3 is the shallower operand pipeline?  (XL/1 shapes, isolated)
for rep in 1 2; do
for cfg in "0 0" "0 4" "0 3" "1 0"; do
  set -- $cfg
  echo "== TMA_EPI=$1 MAX_STAGES=$2 (rep $rep)"
  AB_MODELS=xl VAVAE_GEMM_TMA_EPI=$1 VAVAE_GEMM_MAX_STAGES=$2 timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "w3\|proj\|w12"
done; done
