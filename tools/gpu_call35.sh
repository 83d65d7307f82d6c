#!/bin/bash
# in-situ per-kernel numbers of the north-star sampler (CUDA events around every launch of one kernel kind per run)
for kern in gemm_qkv attention gemm_proj gemm_w12 gemm_w3; do
  timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 2 --no-cpu-baseline --profile-kernel $kern 2>&1 | tail -1 >> gpurun_out/r2_insitu_v2.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r2_insitu_v2.jsonl'):
    d=json.loads(l); r=d['roofline']; print(r['kernel'], round(r['achieved'],1), r['unit'], 'launches', r['launches_timed'], 'step TFLOP/s', round(d['config']['tflops_per_gpu_whole_step'],1), d['clocks']['sm_mhz'])
PY
