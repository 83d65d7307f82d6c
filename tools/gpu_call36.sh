#!/bin/bash
# does the TMA epilogue of the producers change what the consumers find in L2?  in-situ QKV / w12 with and without
for v in 0 2; do for kern in gemm_qkv gemm_w12; do
  echo "== VAVAE_GEMM_TMA_EPI=$v $kern"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 2 --no-cpu-baseline --profile-kernel $kern 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(r['kernel'], round(r['achieved'],1), 'step', round(d['config']['tflops_per_gpu_whole_step'],1), d['clocks']['sm_mhz'])"
done; done
