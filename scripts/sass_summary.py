"""Per-kernel counts of the Blackwell-specific SASS instructions in libromab200.so (no GPU needed):
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load, UTMASTG = TMA store, UTMAREDG = TMA reduce-add,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be absent)."""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "roma_b200", "lib", "libromab200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = {"UTC*MMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA", "LDTM (tcgen05.ld)": r"\bLDTM", "STTM (tcgen05.st)": r"\bSTTM", "UTMALDG (TMA load)": r"\bUTMALDG",
       "UTMASTG (TMA store)": r"\bUTMASTG", "UTMAREDG (TMA reduce)": r"\bUTMAREDG", "UTCBAR (tcgen05.commit)": r"\bUTCBAR",
       "SYNCS (mbarrier)": r"\bSYNCS", "FFMA2 (packed fp32 FMA)": r"\bFFMA2", "HMMA (legacy mma.sync)": r"\bHMMA"}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("void rb::", "").replace("rb::", "")
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pat.items():
            if re.search(p, line):
                counts[cur][k] += 1
print("arch:", sorted(set(re.findall(r"arch = (sm_\w+)", out))))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
print("library totals:", dict(tot))
print()
for k, c in counts.items():
    sel = {n: v for n, v in c.items() if n not in ("SYNCS (mbarrier)", "FFMA2 (packed fp32 FMA)")}
    if sel:
        print(f"{k}\n    " + ", ".join(f"{n}: {v}" for n, v in c.items()))
