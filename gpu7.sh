python -m pytest tests -m gpu -x -q -s 2>&1 | tail -12
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1_c.json; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_r1_c.json'))
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['clocks'], j.get('cpu_baseline'))
i=j['incr_step']; s=i.pop('solve_small_case1'); print(i); print(s)
PY
