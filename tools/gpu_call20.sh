#!/bin/bash
# gated-residual epilogue with TMA-moved residual tiles: correctness, then same-box A/B against the register-prefetch path
echo "== epilogue tests"
timeout -s KILL 300 python -m pytest tests/test_gpu_epilogues.py -m gpu -q -x 2>&1 | tail -5
echo "== parity tests"
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_GEMM_TMA_EPI=$v (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep -v "w12\|qkv"
done; done
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_GEMM_TMA_EPI=$v bench (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
done; done
