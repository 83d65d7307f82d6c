#!/bin/bash
# Final measurements of round 2 (second session build): the driver's command line (10 timed steps), other BASELINE
# configurations, small-batch sweep, in-situ per-kernel numbers, ncu launch list and full captures of the changed kernels.
set -x
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 5 > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
tail -c 2500 gpurun_out/r2b_bench_default.json
timeout -s KILL 300 python bench.py --workload b1 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2b_bench_b1.json 2> gpurun_out/r2b_bench_b1.err
timeout -s KILL 300 python bench.py --workload xl512 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2b_bench_xl512.json 2> gpurun_out/r2b_bench_xl512.err
timeout -s KILL 300 python tools/microbench_forward.py --batches 1,2,4,8,16,32,64 --iters 20 > gpurun_out/r2b_small_batch.jsonl 2>&1
timeout -s KILL 200 python tools/ab_attn.py > gpurun_out/r2b_attn_isolated.log 2>&1
timeout -s KILL 200 python tools/ab_gemms.py > gpurun_out/r2b_gemms_isolated.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r2b_launch_list_xl1.csv \
  python bench.py --points 3 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2b_ncu_ll.log 2>&1
echo "ncu launch list rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_attention_tc|k_gemm_tc" -s 45 -c 5 -o gpurun_out/r2b_xl1_layer \
  python bench.py --points 3 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2b_ncu_full.log 2>&1
echo "ncu full rc=$?"
