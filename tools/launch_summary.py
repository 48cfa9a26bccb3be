"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_summary.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
names = rows[hdr]; ki = names.index('Kernel Name'); vi = names.index('Metric Value'); ui = names.index('Metric Unit')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u in ('ns', 'nsecond') else v
    tot[r[ki][:70]] += v; cnt[r[ki][:70]] += 1
total = sum(tot.values())
print("kernel,launches,total_us,share,us_each")
for k, v in tot.most_common(20): print("%s,%d,%.1f,%.4f,%.1f" % (k.replace(",", ";"), cnt[k], v, v / total, v / cnt[k]))
