#!/bin/bash
# QKV: RoPE table at the front of shared memory / parameter test (cur) vs previous commit (vC.so); interleaved
echo "== qkv epilogue tests"
timeout -s KILL 300 python -m pytest tests/test_gpu_epilogues.py -m gpu -q -x -k qkv 2>&1 | tail -2
for rep in 1 2 3; do for v in vC cur; do
  lib=$PWD/tools/ab/$v.so; [ $v = cur ] && lib=$PWD/va_vae_b200/libvavae_b200.so
  echo "== $v (rep $rep)"
  VAVAE_B200_LIB=$lib timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "qkv"
done; done
for rep in 1 2; do for v in vC cur; do
  lib=$PWD/tools/ab/$v.so; [ $v = cur ] && lib=$PWD/va_vae_b200/libvavae_b200.so
  echo "== $v bench (rep $rep)"
  VAVAE_B200_LIB=$lib timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
