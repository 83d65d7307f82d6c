#!/bin/bash
# ncu capture (source-level stall sampling) of the attention kernel, XL/1 shape, new default (early P V + truncating pack)
AB_MODELS=xl timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep attention
AB_MODELS=xl timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_attention_tc -s 30 -c 1 \
  -o gpurun_out/r2_attn_early python tools/ab_gemms.py > gpurun_out/r2_attn_early_ncu.log 2>&1
echo "ncu rc=$?"
