#!/bin/bash
for rep in 1 2; do for v in 0 2 1; do
  echo "== VAVAE_GEMM_TMA_EPI=$v bench (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
done; done
for v in 0 1; do for kern in gemm_proj gemm_w3; do
  echo "== VAVAE_GEMM_TMA_EPI=$v in-situ $kern"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 2 --no-cpu-baseline --profile-kernel $kern 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline'])"
done; done
