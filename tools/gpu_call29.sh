#!/bin/bash
echo "== epilogue tests"
timeout -s KILL 400 python -m pytest tests/test_gpu_epilogues.py -m gpu -q -x 2>&1 | tail -15
echo "== parity tests"
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
for rep in 1 2; do
  echo "== bench (rep $rep)"
  timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done
timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "proj\|w3\|qkv\|w12"
