"""Per-kernel SASS mnemonic histogram of the built library (no GPU needed): the evidence that the tensor path is tcgen05 / TMA
and that the decode step does not spill.   python tools/sass_summary.py > profiles/r02_sass_mnemonics.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "trace_b200", "libtrace_b200.so")], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
keys = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDSM", "USETMAXREG", "LDL", "STL", "MUFU"]
print("# SASS mnemonic counts per kernel of trace_b200/libtrace_b200.so (`cuobjdump -sass`, sm_100a; tools/sass_summary.py)\n")
print("UTCHMMA = tcgen05.mma (`.2CTA`: cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk (1-D TMA), LDTM / STTM ="
      " tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, HMMA = mma.sync, LDSM = ldmatrix, USETMAXREG = setmaxnreg,"
      " LDL / STL = local-memory (spill) traffic.\n")
print("| kernel | instructions | " + " | ".join(keys) + " |")
print("|---|---|" + "|".join(["---"] * len(keys)) + "|")
for k, c in hist.items():
    row = []
    for key in keys:
        if key == "UTCHMMA":
            row.append(sum(v for n, v in c.items() if n.startswith("UTCHMMA") and ".2CTA" not in n))
        elif key == "UTCHMMA.2CTA":
            row.append(sum(v for n, v in c.items() if n.startswith("UTCHMMA") and ".2CTA" in n))
        else:
            row.append(sum(v for n, v in c.items() if n == key or n.startswith(key + ".")))
    print(f"| `{k}` | {sum(c.values())} | " + " | ".join(str(x) for x in row) + " |")
