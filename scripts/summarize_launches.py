"""profiles/: per-kernel summary + DRAM traffic of one step from the raw ncu launch list
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...)."""
import collections, csv, json, re, sys
raw, out_csv, out_json = sys.argv[1:4]
rows = [r for r in csv.reader(open(raw)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
UNITS = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1,
         'msecond': 1e3, 'second': 1e6}
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[ix['ID']], {'name': re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '')})
    d[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', '')) * UNITS.get(r[ix['Metric Unit']], 1)
agg = collections.OrderedDict()
tot_us = tot_b = ig_b = ig_us = 0.0
for d in per.values():
    a = agg.setdefault(d['name'], [0, 0.0, 0.0])
    b = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    a[0] += 1; a[1] += d['gpu__time_duration.sum']; a[2] += b
    tot_us += d['gpu__time_duration.sum']; tot_b += b
    if 'igemm' in d['name']:
        ig_b += b; ig_us += d['gpu__time_duration.sum']
with open(out_csv, 'w') as f:
    f.write('kernel,launches,total_us,share_of_step,dram_read_plus_write_MB\n')
    for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{n},{t:.1f},{t / tot_us:.4f},{b / 1e6:.1f}\n')
    f.write(f'"TOTAL",{len(per)},{tot_us:.1f},1.0,{tot_b / 1e6:.1f}\n')
json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                     "--profile-from-start off python scripts/profile_step.py --autotune 1 (one forward+explain step, RN50 "
                     "batch 256 bf16); summary: " + out_csv + ", raw: " + raw,
           "igemm_dram_bytes_per_step": ig_b, "all_dram_bytes_per_step": tot_b,
           "igemm_launches": sum(v[0] for k, v in agg.items() if 'igemm' in k), "igemm_share_of_step": ig_us / tot_us,
           "step_us_serialised": tot_us}, open(out_json, 'w'), indent=1)
print(open(out_csv).read())
