#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_epilogues.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest4.log
tail -5 gpurun_out/r2_pytest4.log
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_ROPE_SMEM=$v (rep $rep)"
  VAVAE_ROPE_SMEM=$v timeout -s KILL 200 python tools/ab_gemms.py 2>&1 | grep qkv
done; done 2>&1 | tee gpurun_out/r2_ab_rope.log
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_ROPE_SMEM=$v bench (rep $rep)"
  VAVAE_ROPE_SMEM=$v timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
done; done 2>&1 | tee -a gpurun_out/r2_ab_rope.log
