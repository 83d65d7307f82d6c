#!/bin/bash
# attention: output through a shared-memory tile + TMA store (tma.so) vs direct row-strided stores (base.so)
echo "== correctness (tma)"
timeout -s KILL 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -5
for rep in 1 2; do for v in base tma; do
  echo "== $v (rep $rep)"
  VAVAE_B200_LIB=$PWD/tools/ab/$v.so timeout -s KILL 200 python tools/ab_attn.py 2>&1 | grep attention
done; done
