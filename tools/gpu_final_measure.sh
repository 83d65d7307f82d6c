#!/bin/bash
# Final measurements of a round: the driver's own command line, in-situ per-kernel numbers, the ncu launch list.
mkdir -p gpurun_out
S=$(date +%s)
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s" | tee gpurun_out/r2_bench_default.wall
cat gpurun_out/r2_bench_default.json
: > gpurun_out/r2_insitu_final.jsonl
for k in gemm_qkv attention gemm_proj gemm_w12 gemm_w3; do
  python bench.py --points 16 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --profile-kernel $k >> gpurun_out/r2_insitu_final.jsonl 2>> gpurun_out/r2_insitu_final.err
done
python bench.py --points 3 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r2_launch_list_xl1.csv \
    python bench.py --points 3 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_ncu_ll.log 2>&1
echo "ncu rc=$?"
