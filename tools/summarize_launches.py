"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, avg us, share."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    n = re.sub(r"\(.*", "", r[ki])
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | avg us | share |\n|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{n[:70]}` | {c} | {t/c/1000:.2f} | {t/tot*100:.1f}% |")
