#!/bin/bash
# same-box before / after of the second r2 session: library of commit 8bcdece (tools/ab/r2a.so) vs the current build
for rep in 1 2 3; do for v in r2a cur; do
  lib=$PWD/tools/ab/$v.so; [ $v = cur ] && lib=$PWD/va_vae_b200/libvavae_b200.so
  echo "== $v bench (rep $rep)"
  VAVAE_B200_LIB=$lib timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
for v in r2a cur; do
  lib=$PWD/tools/ab/$v.so; [ $v = cur ] && lib=$PWD/va_vae_b200/libvavae_b200.so
  echo "== $v B/1"
  VAVAE_B200_LIB=$lib timeout -s KILL 200 python bench.py --workload b1 --points 16 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
  echo "== $v xl512"
  VAVAE_B200_LIB=$lib timeout -s KILL 200 python bench.py --workload xl512 --points 16 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done
