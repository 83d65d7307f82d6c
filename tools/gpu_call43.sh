#!/bin/bash
# attention: a quarter of the probabilities through the FMA-pipe polynomial, compiled in (tools/ab/poly.so) vs the committed kernel
echo "== correctness (poly)"
VAVAE_B200_LIB=$PWD/tools/ab/poly.so timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" 2>&1 | tail -2
for rep in 1 2; do for v in cur poly; do
  lib=$PWD/tools/ab/$v.so; [ $v = cur ] && lib=$PWD/va_vae_b200/libvavae_b200.so
  echo "== $v (rep $rep)"
  VAVAE_B200_LIB=$lib timeout -s KILL 200 python tools/ab_attn.py 2>&1 | grep attention
done; done
