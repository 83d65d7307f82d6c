set -x
for c in 48 74 148 296 512; do
  timeout -s KILL 300 python bench.py --points 16 --steps 1 --warmup 3 --no-cpu-baseline --chunk $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B1 chunk',$c, d['value'], d.get('roofline'))"
done
for c in 74 148 256; do
  timeout -s KILL 300 python bench.py --workload xl1 --points 16 --steps 1 --warmup 3 --no-cpu-baseline --chunk $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('XL1 chunk',$c, d['value'], d.get('roofline'))"
done
