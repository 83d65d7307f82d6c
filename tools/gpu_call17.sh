#!/bin/bash
# attention drain: back-to-back stores from distinct registers (default) vs pack-store loop (bit 16)
for v in 0 16; do
  echo "== correctness VAVAE_ATTN_EXPERIMENT=$v"
  VAVAE_ATTN_EXPERIMENT=$v timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" 2>&1 | tail -2
done
for rep in 1 2; do for v in 0 16; do
  echo "== VAVAE_ATTN_EXPERIMENT=$v (rep $rep)"
  VAVAE_ATTN_EXPERIMENT=$v timeout -s KILL 200 python tools/ab_attn.py 2>&1 | grep attention
done; done
