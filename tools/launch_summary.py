"""Summarise an `ncu --csv` launch list by kernel class: python tools/launch_summary.py file.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ik, im, iv, iid = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
per = {}
for r in data:
    if len(r) <= iv: continue
    try: per.setdefault(r[iid], {'name': r[ik]})[r[im]] = float(r[iv].replace(',', ''))
    except ValueError: pass
tot = sum(v.get('gpu__time_duration.sum', 0) for v in per.values())
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for v in per.values():
    name = re.sub(r'^void ', '', v['name']).replace('rchem::', '').replace('(int)', '')
    m = re.search(r'^\w+(<[^>]*>)?', name)
    key = m.group(0) if m else name[:48]
    a = agg[key]; a['n'] += 1
    for k, x in v.items():
        if k == 'name': continue
        if 'registers' in k or 'pct' in k: a[k] = max(a[k], x)
        else: a[k] += x
print(f"total {tot/1e6:.3f} ms over {len(per)} launches")
for k, a in sorted(agg.items(), key=lambda x: -x[1]['gpu__time_duration.sum']):
    t = a['gpu__time_duration.sum']
    extra = ' '.join(f"{kk.split('.')[0][-28:]}={vv:.3g}" for kk, vv in a.items() if kk not in ('n', 'gpu__time_duration.sum'))
    print(f"{k:48s} {t/1e6:9.3f} ms {100*t/tot:5.1f}% n={int(a['n']):3d} {extra}")
