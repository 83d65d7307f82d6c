#!/bin/bash
# TMA epilogue: per-column vectors of the NEXT tile prefetched into registers (cur) vs loaded at the tile start (vB.so)
echo "== epilogue + parity tests"
timeout -s KILL 600 python -m pytest tests/test_gpu_epilogues.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do for v in vB cur; do
  lib=$PWD/tools/ab/$v.so; [ $v = cur ] && lib=$PWD/va_vae_b200/libvavae_b200.so
  echo "== $v (rep $rep)"
  VAVAE_B200_LIB=$lib timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "proj\|w3"
done; done
for rep in 1 2 3; do for v in vB cur; do
  lib=$PWD/tools/ab/$v.so; [ $v = cur ] && lib=$PWD/va_vae_b200/libvavae_b200.so
  echo "== $v bench (rep $rep)"
  VAVAE_B200_LIB=$lib timeout -s KILL 200 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
