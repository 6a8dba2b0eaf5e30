"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source sass` (stdin), with the
largest stall-reason columns of each.  Usage: ... | python tools/ncu_hot_sass.py [N]"""
import csv, sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
def val(r, i):
    try: return float(r[i].replace(",", ""))
    except Exception: return 0.0
for k, hi in enumerate(hdr_i):
    h = rows[hi]
    body = [r for r in rows[hi + 1: hdr_i[k + 1] - 1 if k + 1 < len(hdr_i) else len(rows)] if len(r) == len(h)]
    si = h.index("Warp Stall Sampling (All Samples)")
    ie = h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and not c.endswith("_not_issued")]
    tot = sum(val(r, si) for r in body) or 1.0
    print(f"== {rows[hi - 1][1][:90] if hi else ''}: {len(body)} instructions, {tot:.0f} samples, {sum(val(r, ie) for r in body):.0f} warp instructions")
    agg = {}
    for r in body:
        for i in stall_cols:
            agg[h[i]] = agg.get(h[i], 0.0) + val(r, i)
    print("   stall totals:", ", ".join(f"{a}={b / tot * 100:.1f}%" for a, b in sorted(agg.items(), key=lambda t: -t[1])[:9]))
    idx = sorted(range(len(body)), key=lambda j: -val(body[j], si))[:n]
    for j in sorted(idx):
        r = body[j]
        reasons = sorted(((val(r, i), h[i]) for i in stall_cols if val(r, i) > 0), reverse=True)[:3]
        print(f"{j:5d} {val(r, si) / tot * 100:5.1f}% x{val(r, ie):9.0f} {r[1][:64]:64s} " + " ".join(f"{b[6:]}={a:.0f}" for a, b in reasons))
