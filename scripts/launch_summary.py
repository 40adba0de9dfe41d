"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (kernel, launches,
total ms, average ms, share).  Usage: python scripts/launch_summary.py launches.csv "title / command" > out.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 6 and r[0].isdigit()]
title = sys.argv[2] if len(sys.argv) > 2 else ""
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*$", "", r[4]).strip()
    name = name if len(name) < 110 else name[:107] + "..."
    ns = float(r[-1].replace(",", ""))
    unit = r[-2]
    ms = ns / 1e6 if unit in ("ns", "nsecond") else ns / 1e3 if unit in ("us", "usecond") else ns
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n{title}\n")
print("Cold-cache, serialised per-launch times: compare shares, not absolutes.\n")
print("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
for k, v in tot.most_common():
    print(f"| `{k}` | {cnt[k]} | {v:.3f} | {v / cnt[k]:.4f} | {v / total:.3f} |")
print(f"\nTotal {total:.1f} ms over {sum(cnt.values())} launches.")
