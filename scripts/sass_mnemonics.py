"""Count the SASS mnemonics that identify the Blackwell paths (tcgen05: UTCHMMA / LDTM / STTM / UTCBAR, TMA: UTMALDG,
legacy mma.sync: HMMA, cp.async: LDGSTS, vector reductions: REDG) per kernel of the built library -- what
`cuobjdump -sass` shows, tabulated.  Runs without a GPU.
Usage: python scripts/sass_mnemonics.py [path/to/lib.so] > profiles/<name>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vit-stability-neurodegeneration_b200", "libvsn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = {"UTCHMMA": r"\bUTCHMMA", "UTCHMMA.2CTA": r"UTCHMMA\.2CTA", "UTMALDG": r"\bUTMALDG", "LDTM": r"\bLDTM",
       "STTM": r"\bSTTM", "UTCBAR": r"\bUTCBAR", "HMMA (mma.sync)": r"\bHMMA\.", "LDGSTS (cp.async)": r"\bLDGSTS",
       "REDG": r"\bREDG", "SYNCS (mbarrier)": r"\bSYNCS", "MUFU.EX2": r"MUFU\.EX2"}
cnt, cur = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is not None:
        for k, p in PAT.items():
            if re.search(p, line):
                cnt[cur][k] += 1
names = list(cnt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
fam, nfun = collections.defaultdict(collections.Counter), collections.Counter()
for f, d in zip(names, dem):
    base = re.sub(r"\(.*", "", re.sub(r"<.*", "", d).replace("void ", "").replace("(anonymous namespace)::", ""))
    fam[base].update(cnt[f])
    nfun[base] += 1
keys = list(PAT)
print(f"# SASS mnemonics per kernel of `{os.path.basename(lib)}` (`cuobjdump -sass`, all template instantiations summed)\n")
print("UTCHMMA = `tcgen05.mma` (`.2CTA` = `cta_group::2`), LDTM / STTM = `tcgen05.ld` / `.st`, UTCBAR = `tcgen05.commit`, "
      "UTMALDG = TMA tensor load; HMMA = `mma.sync` (only the fallback kernels of `attn.cu`: windows other than (6,7,6), "
      "head dims other than 32 / 64).\n")
print("| kernel | instantiations | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for b in sorted(fam, key=lambda b: (-fam[b]["UTCHMMA"], -fam[b]["HMMA (mma.sync)"], b)):
    print(f"| `{b}` | {nfun[b]} | " + " | ".join(str(fam[b][k]) for k in keys) + " |")
tot = collections.Counter()
for b in fam:
    tot.update(fam[b])
print("| **total** | | " + " | ".join(str(tot[k]) for k in keys) + " |")
