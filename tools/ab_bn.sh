#!/bin/bash
# Same-box A/B of the GEMM column-tile policy (VAVAE_GEMM_BN): isolated layer kernels, then a 16-point sampler bench.
#   legacy = one width per problem (XL: QKV / proj / w3 run 192-wide tiles), "" = 256-wide + narrower last tile (default)
for rep in 1 2; do
  for v in legacy "" 256 192 128; do
    echo "== VAVAE_GEMM_BN='$v' (rep $rep)"
    VAVAE_GEMM_BN=$v AB_MODELS=${AB_MODELS:-xl} timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | tail -6
  done
done
for rep in 1 2; do
  for v in legacy ""; do
    echo "== VAVAE_GEMM_BN='$v' bench (rep $rep)"
    VAVAE_GEMM_BN=$v timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
  done
done
