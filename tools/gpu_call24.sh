#!/bin/bash
# register-prefetch epilogue: L1 no-allocate residual loads (GEMM_EXPERIMENT=8) x operand-pipeline depth; isolated, XL/1
for ex in 0 8; do for st in 4 5 6; do
  echo "== GEMM_EXPERIMENT=$ex PROJ_STAGES=$st"
  AB_MODELS=xl VAVAE_GEMM_TMA_EPI=0 VAVAE_GEMM_EXPERIMENT=$ex VAVAE_GEMM_PROJ_STAGES=$st timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "proj\|w3"
done; done
for rep in 1 2; do
for cfg in "2 0 0" "0 5 0" "0 5 8" "0 0 8"; do
  set -- $cfg
  echo "== TMA_EPI=$1 PROJ_STAGES=$2 EXPERIMENT=$3 bench (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$1 VAVAE_GEMM_PROJ_STAGES=$2 VAVAE_GEMM_EXPERIMENT=$3 timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
