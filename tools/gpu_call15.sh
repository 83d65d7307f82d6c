#!/bin/bash
# attention: truncating P pack (bit 4) and early P V issue (bit 8), same-box A/B + correctness under each setting
for v in 0 4 8 12; do
  echo "== correctness VAVAE_ATTN_EXPERIMENT=$v"
  VAVAE_ATTN_EXPERIMENT=$v timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" 2>&1 | tail -3
done
for rep in 1 2; do for v in 0 4 8 12; do
  echo "== VAVAE_ATTN_EXPERIMENT=$v (rep $rep)"
  VAVAE_ATTN_EXPERIMENT=$v timeout -s KILL 200 python tools/ab_attn.py 2>&1 | grep attention
done; done
