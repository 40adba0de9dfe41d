"""Aggregate an `ncu --page source --csv --print-source sass` dump by opcode: executed warp instructions and
stall samples.  Usage: ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv; python ncu_sass_mix.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kernel, hdr = None, None
cnt, st, tot = collections.Counter(), collections.Counter(), 0
KEEP = ('MUFU', 'F2FP', 'LDS', 'STS', 'LDL', 'STL', 'LDTM', 'STTM', 'SYNCS', 'BAR', 'F2F', 'I2F', 'F2I', 'LDG', 'STG')


def flush():
    if kernel and tot:
        print(kernel, "total warp instrs", tot)
        for k, v in cnt.most_common(top):
            print(f"  {k:28s} {v:11d} {100 * v / tot:5.1f}%  stall samples {st[k]}")


for r in rows:
    if r and r[0] == "Kernel Name":
        flush()
        kernel, hdr = r[1], None
        cnt, st, tot = collections.Counter(), collections.Counter(), 0
        continue
    if r and r[0] == "Address":
        hdr = r
        ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= ia:
        continue
    op = r[isrc].split()
    if not op:
        continue
    o = op[1] if op[0].startswith('@') else op[0]
    o = o.rstrip(';')
    base = o.split('.')[0]
    key = o if base in KEEP else base
    n = int(r[ia])
    cnt[key] += n
    st[key] += int(r[ist])
    tot += n
flush()
