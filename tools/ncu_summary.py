"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr, agg, tot = None, collections.OrderedDict(), 0.0
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = re.sub(r'\(.*', '', d['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'nbco::\(anonymous namespace\)::|nbco::<unnamed>::', '', name)
    v = float(d['Metric Value'].replace(',', '')); unit = d['Metric Unit']
    ms = v / 1e6 if unit.startswith('n') else (v / 1e3 if unit.startswith('u') else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
print(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"| {k[:80]} | {c} | {ms:.3f} | {ms/tot*100:.1f} % |")
