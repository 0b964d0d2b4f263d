#!/usr/bin/env python
"""Per-kernel SASS digest of rajaperf_b200/lib/librpb200.so (run here, no GPU):  python tools/sass_digest.py [out.md]

`cuobjdump -sass` split per kernel, counting the mnemonics that show which hardware path a kernel uses: UTMALDG (TMA tensor
loads), UBLKCP (bulk-async copies), DMMA (FP64 tensor-core MMA), 256-bit LDG / STG, SYNCS (mbarrier), ATOMS / ATOMG / RED,
VOTE / SHFL / MATCH, BAR, MEMBAR.SYS and system-scope LD / ST (the halo flags), plus the instruction count and the target."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rajaperf_b200", "lib", "librpb200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_digest.md")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
try:
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
except Exception:
    commit = "?"

COLS = [("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("DMMA", r"\bDMMA"), ("LDG.256", r"\bLDG\S*\.256"), ("STG.256", r"\bSTG\S*\.256"),
        ("LDG.128", r"\bLDG\S*\.128"), ("STG.128", r"\bSTG\S*\.128"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("BAR", r"\bBAR\."),
        ("VOTE", r"\bVOTE"), ("SHFL", r"\bSHFL"), ("ATOMS", r"\bATOMS"), ("ATOMG/RED", r"\b(ATOMG|RED)\b"),
        ("DFMA/DADD/DMUL", r"\b(DFMA|DADD|DMUL)\b"), ("LD/ST .SYS", r"\b(LD|ST|LDG|STG)\S*\.SYS"), ("MEMBAR.SYS", r"MEMBAR\S*\.SYS"),
        ("local (spill)", r"\b(LDL|STL)\b")]
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        kernels[cur].append(line)


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        clean = lambda o: re.sub(r"\((bool|int|unsigned int|unsigned)\)", "", o.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""))
        return [re.sub(r"\(.*", "", clean(o)) for o in out]
    except Exception:
        return names


names = demangle(list(kernels))
rows = []
for (mangled, lines), name in zip(kernels.items(), names):
    text = "\n".join(lines)
    rows.append((name, len(lines), [len(re.findall(p, text)) for _, p in COLS]))
rows.sort(key=lambda r: r[0])
L = [f"# SASS digest of `rajaperf_b200/lib/librpb200.so` at {commit} (tools/sass_digest.py)", "",
     f"targets in the fat binary: {', '.join(archs)} -- sm_100a only; {len(rows)} kernels.  Counts are static instruction counts.", "",
     "| kernel | instr | " + " | ".join(c for c, _ in COLS) + " |", "|---|---|" + "---|" * len(COLS)]
for name, n, counts in rows:
    L.append(f"| `{name[:110]}` | {n} | " + " | ".join(str(c) if c else "" for c in counts) + " |")
tot = [sum(r[2][i] for r in rows) for i in range(len(COLS))]
L += ["", "Totals: " + ", ".join(f"{c} {t}" for (c, _), t in zip(COLS, tot) if t),
      "", "No `UTCMMA` / `LDTM` (tcgen05) anywhere: every kernel on the path is FP64 or byte/index work, and tcgen05 has no FP64 kind.",
      f"tcgen05 mnemonics found: {len(re.findall(r'UTCMMA|LDTM|UTCBAR', sass))}."]
open(out_path, "w").write("\n".join(L) + "\n")
print("\n".join(L[:4]))
print(L[-3])
