"""Aggregate an ncu launch list (gpu__time_duration.sum CSV) per kernel name."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
kn, mv, mn, mu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Name'), H.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if r[mn] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[kn]).replace('<unnamed>::', '').replace('void ', '')
    name = re.sub(r'<.*', '', name)
    v = float(r[mv].replace(',', ''))
    v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1.0}.get(r[mu], 1.0)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
title = sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"
print(f"# {title}\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"| `{k[:60]}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
print(f"\ntotal {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
