"""Per-kernel SASS mnemonic counts of build/*.o (cuobjdump -sass) -> profiles/r2_sass_grep.md: evidence that the TMA / vector-reduction /
async-copy paths named in DESIGN.md are really in the binary, and how many FFMA each kernel holds (the decision arithmetic of the
exact-semantics kernels is written with __fmul_rn/__fadd_rn, which ptxas never contracts; FFMAs there come from the conservative pruning
tests around it).  Usage: python tools/sass_grep.py > profiles/r2_sass_grep.md"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("RED/REDG .F32x4", r"\bRED[G]?\.[A-Z0-9.]*F32x4|\bREDG\.E\.ADD\.F32x4|\bRED\.E\.ADD\.F32x4"),
       ("RED (all)", r"\bREDG?\b|\bRED\."), ("ATOM(G/S)", r"\bATOM[GS]?\b|\bATOMG\.|\bATOMS\."), ("MATCH", r"\bMATCH\."), ("SHFL", r"\bSHFL\."),
       ("REDUX", r"\bREDUX"), ("FFMA", r"\bFFMA"), ("FMUL", r"\bFMUL"), ("FADD", r"\bFADD"), ("MUFU", r"\bMUFU"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
       ("LDG", r"\bLDG"), ("STG", r"\bSTG")]
print("# SASS mnemonic counts per kernel (sm_100a, `cuobjdump -sass build/*.o`)\n")
print("Counts are static instruction counts in the kernel body.  `UBLKCP` = 1-D TMA bulk copy (cp.async.bulk), `SYNCS` = mbarrier ops,")
print("`LDGSTS` = cp.async, `RED .F32x4` = 16-byte vector reduction, `MATCH` = __match_any_sync, `REDUX` = __reduce_*_sync.\n")
print("| object | kernel | " + " | ".join(n for n, _ in PAT) + " |")
print("|---|---|" + "---|" * len(PAT))
for obj in sorted(glob.glob(os.path.join(ROOT, "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = [0] * len(PAT)
            continue
        if cur is None or "/*" not in line:
            continue
        for i, (_, pat) in enumerate(PAT):
            if re.search(pat, line):
                counts[cur][i] += 1
    for k, c in counts.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("dtb::", "")
        print("| %s | `%s` | %s |" % (os.path.basename(obj), name[:70], " | ".join(str(x) for x in c)))
