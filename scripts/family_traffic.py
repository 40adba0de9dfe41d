"""DRAM traffic per kernel family of ONE optimiser step, from an ncu launch list taken with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
over `bench.py` (graph mode: the kernels of the replayed graphs are profiled as nodes).  Steps are delimited by the fused
AdamW launch (one per optimiser step); the LAST complete step of the list is summarised, so warm-up, capture and
rehearsal passes do not count.  Prints a markdown table and merges `family:<name>` entries into profiles/ncu_traffic.json
(read by bench.py for `roofline.traffic`).
Usage: python scripts/family_traffic.py launches.csv "command line" profiles/ncu_traffic.json > profiles/<name>.md"""
import collections
import csv
import json
import re
import sys


def family(kernel: str) -> str:
    if "gemm_tc_kernel" in kernel:
        return "gemm"
    if "attn_bwd" in kernel or "attn_delta" in kernel or "dattn_dq_cast" in kernel or "bias_table_grad" in kernel:
        return "attention_bwd"
    if "attn_fwd" in kernel:
        return "attention_fwd"
    if re.search(r"\bln_|layernorm|merge_ln|patch_ln|colreduce", kernel):
        return "layernorm"
    return "elementwise_layout_optim"


def to_bytes(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    u = unit.lower()
    return x * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)


def to_ms(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    u = unit.lower()
    return x / 1e6 if u in ("ns", "nsecond") else x / 1e3 if u in ("us", "usecond") else x * 1e3 if u in ("s", "second") else x


def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(r for r in rows if "Kernel Name" in r and "Metric Name" in r)
    i_id, i_k, i_m, i_u, i_v = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()         # ncu ID -> {kernel, ms, rd, wr}
    for r in rows:
        if len(r) <= i_v or not r[i_id].isdigit():
            continue
        e = launches.setdefault(int(r[i_id]), {"kernel": re.sub(r"\(.*$", "", r[i_k]).strip(), "ms": 0.0, "rd": 0.0, "wr": 0.0})
        m = r[i_m]
        if m.startswith("gpu__time_duration"):
            e["ms"] = to_ms(r[i_v], r[i_u])
        elif m.startswith("dram__bytes_read"):
            e["rd"] = to_bytes(r[i_v], r[i_u])
        elif m.startswith("dram__bytes_write"):
            e["wr"] = to_bytes(r[i_v], r[i_u])
    seq = list(launches.values())
    ends = [i for i, e in enumerate(seq) if "mt_adamw_kernel" in e["kernel"]]
    if len(ends) < 2:
        raise SystemExit("family_traffic: fewer than two optimiser steps in the launch list")
    step = seq[ends[-2] + 1: ends[-1] + 1]
    fam = collections.OrderedDict()
    for e in step:
        f = fam.setdefault(family(e["kernel"]), {"launches": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
        f["launches"] += 1
        f["ms"] += e["ms"]
        f["rd"] += e["rd"]
        f["wr"] += e["wr"]
    tot_ms = sum(f["ms"] for f in fam.values())
    print("# DRAM traffic per kernel family of one optimiser step (ncu, --clock-control none)\n")
    print(title + "\n")
    print(f"Last complete step of the list: {len(step)} launches, {tot_ms:.2f} ms of serialised, cold-cache kernel time "
          "(compare shares and bytes, not absolute times).\n")
    print("| family | launches | ms (ncu) | share | DRAM read MB | DRAM write MB | DRAM total MB |\n|---|---|---|---|---|---|---|")
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"| {k} | {f['launches']} | {f['ms']:.3f} | {f['ms'] / tot_ms:.3f} | {f['rd'] / 1e6:.1f} | {f['wr'] / 1e6:.1f} | "
              f"{(f['rd'] + f['wr']) / 1e6:.1f} |")
    print(f"\nWhole step: {sum(f['rd'] + f['wr'] for f in fam.values()) / 1e9:.2f} GB of DRAM traffic.")
    # per kernel inside the step
    ker = collections.OrderedDict()
    for e in step:
        k = ker.setdefault(e["kernel"], {"n": 0, "ms": 0.0, "b": 0.0})
        k["n"] += 1
        k["ms"] += e["ms"]
        k["b"] += e["rd"] + e["wr"]
    print("\n| kernel | launches | ms (ncu) | share | DRAM MB | GB/s under ncu |\n|---|---|---|---|---|---|")
    for name, k in sorted(ker.items(), key=lambda kv: -kv[1]["ms"])[:28]:
        nm = name if len(name) < 90 else name[:87] + "..."
        print(f"| `{nm}` | {k['n']} | {k['ms']:.3f} | {k['ms'] / tot_ms:.3f} | {k['b'] / 1e6:.1f} | "
              f"{k['b'] / (k['ms'] * 1e-3) / 1e9 if k['ms'] else 0:.0f} |")
    if out_json:
        try:
            with open(out_json) as f:
                tr = json.load(f)
        except OSError:
            tr = {}
        for k, f in fam.items():
            tr["family:" + k] = {"dram_bytes_per_step": int(f["rd"] + f["wr"]), "launches_per_step": f["launches"],
                                 "source": f"{title} (scripts/family_traffic.py over the ncu launch list, last complete step)"}
        with open(out_json, "w") as f:
            json.dump(tr, f, indent=1)


if __name__ == "__main__":
    main()
