#!/bin/bash
for rep in 1 2; do for v in 0 32; do
  echo "== VAVAE_ATTN_EXPERIMENT=$v (rep $rep)"
  VAVAE_ATTN_EXPERIMENT=$v timeout -s KILL 200 python tools/ab_attn.py 2>&1 | grep attention
done; done
