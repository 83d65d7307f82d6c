#!/bin/bash
# same-box in-situ QKV / attention before (r2a.so = commit 8bcdece) and after; configs[3] line
for v in r2a cur; do
  lib=$PWD/tools/ab/$v.so; [ $v = cur ] && lib=$PWD/va_vae_b200/libvavae_b200.so
  for kern in gemm_qkv attention gemm_w12; do
  echo "== $v in-situ $kern"
  VAVAE_B200_LIB=$lib timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 2 --no-cpu-baseline --e2e-steps 0 --profile-kernel $kern 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(r['kernel'], round(r['achieved'],1), 'step', round(d['config']['tflops_per_gpu_whole_step'],1), d['clocks']['sm_mhz'])"
  done
  echo "== $v isolated"
  AB_MODELS=xl VAVAE_B200_LIB=$lib timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep "qkv\|w12"
done
timeout -s KILL 300 python bench.py --workload xl1_512 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2b_bench_xl512.json 2> gpurun_out/r2b_bench_xl512.err
tail -c 900 gpurun_out/r2b_bench_xl512.json
