#!/bin/bash
for v in 0 2; do
  echo "== VAVAE_PDL=$v"
  VAVAE_PDL=$v timeout -s KILL 300 python tools/microbench_forward.py --batches 1,2,4,8,16,32,64 --iters 20 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['batch'], round(d['ms_per_forward'], 3), 'ms', round(d['tflops'], 1), 'TFLOP/s')
"
done 2>&1 | tee gpurun_out/r2_ab_pdl_graph.log
VAVAE_PDL=2 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph or sharding" 2>&1 | tail -3
