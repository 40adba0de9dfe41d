"""Summarise an `ncu --set full --import-source on` capture (.ncu-rep) as markdown: duration, pipe utilisation, DRAM
traffic, warp-stall samples, executed-instruction mix and the hottest SASS lines.
Usage: python scripts/ncu_summary.py capture.ncu-rep [launch_index] > profiles/<name>.md"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0


def page(kind):
    out = subprocess.run(["ncu", "-i", rep, "--page", kind, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


raw = page("raw")
hdr, units, row = raw[0], raw[1], raw[2 + which]
get = lambda k: next((row[i] for i, h in enumerate(hdr) if h == k), "n/a")   # noqa: E731
unit = lambda k: next((units[i] for i, h in enumerate(hdr) if h == k), "")   # noqa: E731
print(f"# ncu summary: `{rep.split('/')[-1]}` (launch {which})\n")
print(f"Kernel: `{get('Kernel Name')}`  grid {get('launch__grid_size')} x block {get('launch__block_size')}, "
      f"{get('launch__registers_per_thread')} registers/thread\n")
print("| metric | value |\n|---|---|")
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]:
    tail = k if k in hdr else next((h for h in hdr if h.endswith(k)), None)
    if tail:
        print(f"| `{k}` | {get(tail)} {unit(tail)} |")

src = page("source")
if len(src) > 2:
    h = src[1]
    body = [x for x in src[2:] if len(x) >= len(h)]
    i_s, i_e, i_n = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    agg, mix, lines, tot = collections.Counter(), collections.Counter(), [], 0
    for idx, x in enumerate(body):
        for i in stall:
            if x[i] not in ("0", ""):
                agg[h[i][6:]] += int(x[i])
        toks = x[i_s].split()
        op = (toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else toks[0] if toks else "").split(".")[0]
        e = int(x[i_e] or 0)
        mix[op] += e
        tot += e
        lines.append((int(x[i_n] or 0), idx, x[i_s].strip(), e,
                      {h[i][6:]: int(x[i]) for i in stall if x[i] not in ("0", "")}))
    ns = sum(agg.values())
    print(f"\nWarp-stall samples ({ns}): " + ", ".join(f"{k} {v / ns:.1%}" for k, v in agg.most_common(8)))
    print(f"\nExecuted warp instructions by opcode (source page, {tot}): " +
          ", ".join(f"{k} {v / tot:.1%}" for k, v in mix.most_common(14)))
    print("\n| samples | executed | SASS | top stalls |\n|---|---|---|---|")
    for s, idx, text, e, st in sorted(lines, reverse=True)[:14]:
        top = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"| {s} | {e} | `{text[:70]}` | {top} |")
