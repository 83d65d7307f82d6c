#!/bin/bash
# One 8-GPU box: (1) sharded result == single-GPU result at world 8 (NCCL), (2) STRONG scaling of BASELINE configs[2]
# (XL/1, global batch 1024 sharded over 1/2/4/8 GPUs: 1024/512/256/128 samples per GPU) on a short grid, through
# va_vae_b200.dist.sample_sharded, (3) the default weak-scaling workload at N=8 on the same short grid.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/r2_multi_gpu_check8.log 2>&1; echo "check rc=$?"; tail -2 gpurun_out/r2_multi_gpu_check8.log
: > gpurun_out/r2_strong_scaling.jsonl
for n in 1 2 4 8; do
  if [ $n = 1 ]; then L="python"; else L="$TR --nproc-per-node $n --master-port $((29520+n))"; fi
  $L bench.py --gpus $n --global-n 1024 --points 10 --steps 1 --warmup 1 --warmup-points 4 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/r2_strong_$n.err | tail -1 >> gpurun_out/r2_strong_scaling.jsonl
  echo "strong N=$n rc=$?"
done
$TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --points 10 --steps 1 --warmup 1 --warmup-points 4 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/r2_weak_8.err | tail -1 > gpurun_out/r2_weak8.json
python - <<'PY'
import json
for l in open('gpurun_out/r2_strong_scaling.jsonl'):
    d=json.loads(l); c=d['config']
    print(d['n_gpus'], round(d['value'],2), 'samples/s', round(c['tflops_per_gpu_whole_step'],1), 'TFLOP/s/GPU', c['chunk_rows'], d['clocks']['sm_mhz'], d['ms_per_step'])
d=json.load(open('gpurun_out/r2_weak8.json')); print('weak8', d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])
PY
