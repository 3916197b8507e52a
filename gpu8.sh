python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r1_d.json; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_r1_d.json'))
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['fp32_probe_tflops'])
i=j['incr_step']; s=i.pop('solve_small_case1'); print(i); print(s['s_per_incr_step'])
PY
