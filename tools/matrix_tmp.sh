python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; tail -2 gpurun_out/r2_bench_n1_b.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_n1_b.json'))
print("cfg3", round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 1), round(d['roofline']['frac_executed'], 3))
for k, v in d['other_workloads'].items(): print(k, round(v['value'], 1), round(v['e2e']['value'], 1), round(v['ms_per_step'], 2), round(v.get('frac_executed') or 0, 3), v['succeeded_fraction'])
PY
