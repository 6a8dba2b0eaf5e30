import json, sys
d = json.load(open(sys.argv[1]))
print('sum_ms', d['sum_ms'])
agg = {}
for r in d['per_step']:
    k = r['kernel'].split(' ')[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += r['launches']; a[1] += r['ms']
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print('%-28s %6.1f launches %8.3f ms' % (k, v[0], v[1]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
for r in d['per_step'][:n]:
    print('%-70s %5.1f %8.3f ms %s' % (r['kernel'][:70], r['launches'], r['ms'], ('%.0f TF' % r['tflops']) if r['tflops'] else ''))
