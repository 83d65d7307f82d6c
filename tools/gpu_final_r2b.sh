#!/bin/bash
# Final measurements of round 2 (second-session build): tests, the driver's command line (10 timed steps), the other
# BASELINE configurations, small-batch sweep, in-situ per-kernel numbers, ncu launch list, reference arm.
set -x
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2b_pytest_final.log; cat gpurun_out/r2b_pytest_final.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/r2b_smoke.log; cat gpurun_out/r2b_smoke.log
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 5 > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
tail -c 1200 gpurun_out/r2b_bench_default.json
timeout -s KILL 300 python bench.py --workload b1 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2b_bench_b1.json 2> gpurun_out/r2b_bench_b1.err
timeout -s KILL 300 python bench.py --workload xl1_512 --method euler --cfg-interval-start 0.11 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2b_bench_xl512_skip.json 2> gpurun_out/r2b_bench_xl512.err
timeout -s KILL 300 python bench.py --workload xl1_512 --method euler --cfg-interval-start 0.11 --literal-2n --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2b_bench_xl512_literal.json 2>> gpurun_out/r2b_bench_xl512.err
timeout -s KILL 300 python tools/microbench_forward.py --batches 1,2,4,8,16,32,64 --iters 20 > gpurun_out/r2b_small_batch.jsonl 2>&1
rm -f gpurun_out/r2b_insitu.jsonl
for kern in gemm_qkv attention gemm_proj gemm_w12 gemm_w3; do
  timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 2 --no-cpu-baseline --e2e-steps 0 --profile-kernel $kern 2>/dev/null | tail -1 >> gpurun_out/r2b_insitu.jsonl
done
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r2b_launch_list_xl1.csv \
  python bench.py --points 3 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2b_ncu_ll.log 2>&1
echo "ncu launch list rc=$?"
timeout -s KILL 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err
tail -c 600 gpurun_out/r2b_bench_reference.json
