#!/bin/bash
# two-GPU sanity of the final build: torchrun bench (short grid), the two-GPU single-process test, long-trajectory numbers
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --points 16 --steps 1 --warmup 2 --no-cpu-baseline 2>gpurun_out/r2b_2gpu.err | tail -1 > gpurun_out/r2b_bench_2gpu.json
python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_2gpu.json').read()); print('2 GPUs:', d['value'], d['n_gpus'], d['config']['tflops_per_gpu_whole_step'], d['e2e']['value'], d['clocks'])"
timeout -s KILL 300 python -m pytest tests -m gpu -q -k "two_gpu or second_device or cuda1" 2>&1 | tail -2
timeout -s KILL 300 python -m pytest tests/test_gpu_sampler_extras.py -m gpu -q -s -k "long_trajectory" 2>&1 | grep -i "long trajectory\|passed\|failed" | tee gpurun_out/r2b_long_trajectory.log
