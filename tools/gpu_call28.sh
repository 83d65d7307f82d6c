#!/bin/bash
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_PROJ_WIDE=$v (rep $rep)"
  VAVAE_PROJ_WIDE=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "proj"
done; done
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_PROJ_WIDE=$v bench (rep $rep)"
  VAVAE_PROJ_WIDE=$v timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
