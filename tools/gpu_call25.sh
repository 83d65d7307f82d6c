#!/bin/bash
# batch-1 XL/1 forward: per-kernel durations (ncu launch list, graph nodes profiled individually)
timeout -s KILL 300 python tools/microbench_forward.py --batches 1,2,4,8 --iters 20 2>&1 | cut -c1-200
VAVAE_GRAPH_MAX_ROWS=0 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launch_list_b1fwd.csv python tools/microbench_forward.py --batches 1 --iters 1 > gpurun_out/r2_ncu_b1fwd.log 2>&1
echo "ncu rc=$?"
