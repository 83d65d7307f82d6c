#!/bin/bash
echo "== full GPU suite"
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== smoke"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
for rep in 1 2; do
  echo "== bench (rep $rep)"
  timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done
