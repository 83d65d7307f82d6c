#!/bin/bash
# TMA epilogue: eight warps for attn.proj, four warps (five operand stages) for the XL mlp.w3
for v in 1 2; do
echo "== epilogue tests (VAVAE_GEMM_TMA_EPI=$v)"
VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 300 python -m pytest tests/test_gpu_epilogues.py -m gpu -q -x 2>&1 | tail -2
done
for rep in 1 2; do for v in 0 2; do
  echo "== VAVAE_GEMM_TMA_EPI=$v (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "proj\|w3"
done; done
for rep in 1 2 3; do for v in 0 2; do
  echo "== VAVAE_GEMM_TMA_EPI=$v bench (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
