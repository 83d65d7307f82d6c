#!/bin/bash
# four-warp TMA epilogue (ring depth 4 for proj, 2 for w3 with the fifth operand stage kept)
echo "== epilogue tests (TMA for every eligible launch)"
VAVAE_GEMM_TMA_EPI=1 timeout -s KILL 300 python -m pytest tests/test_gpu_epilogues.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_GEMM_TMA_EPI=$v (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "proj\|w3"
done; done
echo "== TMA_EPI=1 XT_SLOTS=2 / 3"
for ns in 2 3; do AB_MODELS=xl VAVAE_GEMM_TMA_EPI=1 VAVAE_GEMM_XT_SLOTS=$ns timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "proj\|w3"; done
for rep in 1 2; do for v in 0 2 1; do
  echo "== VAVAE_GEMM_TMA_EPI=$v bench (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
