"""Per-kernel SASS evidence that the product library is Blackwell-native: counts of the tcgen05 / TMEM / TMA mnemonics
(UTCHMMA = tcgen05.mma kind::f16/tf32, UTCHMMA.2CTA = cta_group::2, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR =
tcgen05.commit, SYNCS = mbarrier) in every kernel of lib/libqmri_b200.so.
    python profiles/tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "qmri-pnp-recon-poc_b200", "lib", "libqmri_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
MNEMONICS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "SYNCS", "UCGABAR", "R2UR", "ELECT", "FFMA2", "FFMA", "HMMA", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMG", "RED", "SHFL"]
counts = collections.OrderedDict()
arch = None
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["total"] += 1
    for mn in MNEMONICS:
        if mn == "FFMA" and op.startswith("FFMA2"):
            continue
        if op == mn or op.startswith(mn + "."):
            counts[cur][mn] += 1
            if mn == "UTCHMMA.2CTA":
                break
            if mn == "UTCHMMA" and ".2CTA" in op:
                counts[cur]["UTCHMMA.2CTA"] += 1
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(so, ROOT)}: arch {arch}, {len(counts)} kernels; instruction counts per kernel (static SASS)")
cols = [m for m in MNEMONICS if any(c[m] for c in counts.values())]
print("| kernel | total | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
tot = collections.Counter()
for (name, c), dn in zip(counts.items(), demangle):
    short = re.sub(r"\(anonymous namespace\)::", "", dn)
    short = re.sub(r"\(.*$", "", short)
    short = re.sub(r"^void ", "", short)
    print(f"| `{short}` | {c['total']} | " + " | ".join(str(c[m]) if c[m] else "" for m in cols) + " |")
    tot.update(c)
print(f"| **all kernels** | {tot['total']} | " + " | ".join(str(tot[m]) for m in cols) + " |")
