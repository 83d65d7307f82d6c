#!/bin/bash
# four-warp TMA epilogue for every mlp.w3 (default) vs the XL-only rule (VAVAE_GEMM_TMA_EPI=3): B/1 sampler, then the full suite
for rep in 1 2; do for v in 3 2; do
  echo "== B/1 VAVAE_GEMM_TMA_EPI=$v (rep $rep)"
  VAVAE_GEMM_TMA_EPI=$v timeout -s KILL 200 python bench.py --workload b1 --points 16 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks']['sm_mhz'])"
done; done
timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
