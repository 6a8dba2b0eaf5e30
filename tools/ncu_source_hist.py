"""ncu -i X.ncu-rep --page source --csv --kernel-name regex:K ...  ->  opcode histogram + hottest stall lines."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0
ops = collections.Counter()
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Instructions Executed"]] == "Instructions Executed":
        continue
    try:
        n = int(r[ix["Instructions Executed"]]); st = int(r[ix["Warp Stall Sampling (All Samples)"]])
    except ValueError:
        continue
    tot += n
    toks = r[ix["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ops[op.split(".")[0]] += n
    data.append((n, st, r[ix["Source"]].strip()))
print("total inst", tot)
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("  %-12s %12d %5.1f%%" % (k, v, 100.0 * v / tot))
print("--- top stall samples (inst executed, samples, sass)")
for n, st, src in sorted(data, key=lambda x: -x[1])[:30]:
    print("%10d %8d  %s" % (n, st, src[:110]))
