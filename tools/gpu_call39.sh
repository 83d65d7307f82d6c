#!/bin/bash
for rep in 1 2; do for v in r2a vA vB; do
  echo "== $v (rep $rep)"
  AB_MODELS=xl VAVAE_B200_LIB=$PWD/tools/ab/$v.so timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "qkv"
done; done
