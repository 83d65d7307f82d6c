#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest13.log; tail -3 gpurun_out/r2_pytest13.log
python bench.py --points 16 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --profile-kernel gemm_proj 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('proj', d['roofline']['achieved'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
python bench.py --points 16 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('xl1 16pt', d['value'], d['config']['tflops_per_gpu_whole_step'], d['clocks'])"
python bench.py --workload b1 --steps 2 --warmup 3 > gpurun_out/r2_bench_b1.json 2>gpurun_out/r2_bench_b1.err; tail -c 1500 gpurun_out/r2_bench_b1.json | head -c 400; echo
python bench.py --workload xl1_512 --method euler --cfg-interval-start 0.11 --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_xl512_skip.json 2>gpurun_out/r2_bench_xl512.err
python bench.py --workload xl1_512 --method euler --cfg-interval-start 0.11 --literal-2n --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_xl512_literal.json 2>>gpurun_out/r2_bench_xl512.err
for f in skip literal; do python -c "
import json; d=json.load(open('gpurun_out/r2_bench_xl512_$f.json')); c=d['config']; print('$f', d['value'], d['ms_per_step'], c['forwards_executed_per_step'], c['forwards_reference_per_step'], c['tflops_per_gpu_whole_step'])"; done
python bench.py --points 3 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 42 -c 4 -o gpurun_out/r2_xl1_gemms_final \
    python bench.py --points 3 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_ncu_final.log 2>&1
echo "ncu rc=$?"
