#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
bash tools/ab_bn.sh > gpurun_out/r2_ab_bn.log 2>&1
cat gpurun_out/r2_ab_bn.log
