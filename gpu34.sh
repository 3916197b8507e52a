python -m pytest tests/test_flow_gpu.py -x -q 2>&1 | tail -3
python benchmarks/micro_bench.py --quick 2>/dev/null | grep 'train' | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['op'], r['d'], r['n'], round(r['ms'],3),'ms', round(r.get('samples_per_s',0)/1e6,1),'M/s', r.get('us_per_iter'))
"
