#!/bin/bash
# First GPU call of round 2: tests, the new default bench (short), in-situ per-kernel numbers on XL/1, one ncu capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_xl1.json 2> gpurun_out/r2_bench_xl1.err; echo "bench rc=$?"
cat gpurun_out/r2_bench_xl1.json
: > gpurun_out/r2_insitu.jsonl
for k in gemm_qkv attention gemm_proj gemm_w12 gemm_w3 norm_modulate; do
  python bench.py --points 16 --steps 1 --warmup 1 --no-cpu-baseline --profile-kernel $k >> gpurun_out/r2_insitu.jsonl 2>> gpurun_out/r2_insitu.err
done
python bench.py --points 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 42 -c 4 -o gpurun_out/r2_xl1_gemms \
    python bench.py --points 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu.log 2>&1
echo "ncu rc=$?"
