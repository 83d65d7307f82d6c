#!/bin/bash
# Same-box A/B of one environment knob: tools/ab_env.sh VAR script.py   (values 0 and 1, two repetitions, then a
# 16-point bench per value)
var=$1; script=$2
for rep in 1 2; do
  for v in 0 1; do
    echo "== $var=$v (rep $rep)"
    env $var=$v timeout -s KILL 200 python $script 2>&1 | tail -8
  done
done
for rep in 1 2; do
  for v in 0 1; do
    echo "== $var=$v bench (rep $rep)"
    env $var=$v timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
  done
done
