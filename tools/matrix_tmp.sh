python -m pytest tests/test_gpu_hpmf.py -q -x 2>&1 | tail -4
python tools/time_hpmf_fused.py cfg3 256,2048 2>&1 | grep "C="
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hpmf|finite_mask|sasa" -c 4000 --csv --log-file gpurun_out/r2_hpmf_launches.csv python tools/time_hpmf_fused.py cfg3 256 > gpurun_out/ncu_hpmf.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2_hpmf_launches.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
names = rows[hdr]; ki = names.index('Kernel Name'); vi = names.index('Metric Value'); ui = names.index('Metric Unit')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u in ('ns', 'nsecond') else v   # -> us
    tot[r[ki][:60]] += v; cnt[r[ki][:60]] += 1
for k, v in tot.most_common(14): print("%-62s %6d launches %10.1f us  (%.1f us each)" % (k, cnt[k], v, v / cnt[k]))
PY
