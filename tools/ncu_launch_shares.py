"""ncu --metrics gpu__time_duration.sum launch list (csv) -> per-kernel share table (markdown)."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4].replace("void ", "")).replace("eavit::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.3f} ms under ncu (cold-cache, serialised: compare SHARES, not absolute times)\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k[:90]} | {n} | {us:.1f} | {100 * us / tot:.1f} % |")
