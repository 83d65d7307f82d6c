#!/bin/bash
for v in 0 1 2 3; do
  echo "== VAVAE_GEMM_NARROW=$v"
  VAVAE_GEMM_NARROW=$v timeout -s KILL 300 python tools/microbench_forward.py --batches 1,2,4,8,16 --iters 20 2>&1 | cut -c30-110
done
