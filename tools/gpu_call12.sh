#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do for v in 0 1; do
  echo "== VAVAE_ALT_ORDER=$v bench (rep $rep)"
  VAVAE_ALT_ORDER=$v timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
done; done 2>&1 | tee gpurun_out/r2_ab_alt_order.log
for v in 0 1; do
  echo "== VAVAE_ALT_ORDER=$v b1 bench"
  VAVAE_ALT_ORDER=$v timeout -s KILL 300 python bench.py --workload b1 --points 16 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
done 2>&1 | tee -a gpurun_out/r2_ab_alt_order.log
