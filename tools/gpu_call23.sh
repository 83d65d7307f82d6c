#!/bin/bash
# attn.proj: operand-pipeline depth sweep for both epilogues (isolated), then the sampler for the interesting points
for t in 0 2; do for st in 2 3 4 5 6; do
  echo "== TMA_EPI=$t PROJ_STAGES=$st"
  AB_MODELS=xl VAVAE_GEMM_TMA_EPI=$t VAVAE_GEMM_PROJ_STAGES=$st timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "proj"
done; done
for rep in 1 2; do
for cfg in "0 0" "0 4" "0 3" "2 0" "2 3"; do
  set -- $cfg
  echo "== TMA_EPI=$1 PROJ_STAGES=$2 bench (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$1 VAVAE_GEMM_PROJ_STAGES=$2 timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
