"""Print the per-call table written by `bench.py --profile-out`."""
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("instrumented step ms", round(d['step_ms_instrumented'], 2), "sum of calls", round(d['sum_of_calls_ms'], 2))
by = collections.Counter()
for r in d['calls']:
    by[r['call']] += r['ms_total']
for k, v in by.most_common():
    print(f"  {k:24s} {v:7.3f}")
for r in d['calls'][:top]:
    print(f"{r['call']:20s} {r['shape']:46s} n={r['launches']:3d} tot={r['ms_total']:6.3f} per={r['ms_per_call']:.4f} "
          f"tf={r.get('tflops', '')} gbs={r.get('gbs', '')}")
