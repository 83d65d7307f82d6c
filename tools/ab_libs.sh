#!/bin/bash
# Same-box A/B of several builds of the library: tools/ab/<name>.so (built by hand, git-ignored).
# usage: tools/ab_libs.sh <script.py> name1 name2 ...   then a 16-point bench per variant
script=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    echo "== $v (rep $rep)"
    VAVAE_B200_LIB=$PWD/tools/ab/$v.so timeout -s KILL 200 python $script 2>&1 | tail -8
  done
done
for rep in 1 2; do
  for v in "$@"; do
    echo "== $v bench (rep $rep)"
    VAVAE_B200_LIB=$PWD/tools/ab/$v.so timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
  done
done
