"""Opcode histogram of the built library's SASS, per kernel family (evidence for which hardware paths the kernels use:
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA),
UTMALDG / UTMASTG = TMA tensor copies, FFMA2 / FMUL2 = packed fp32x2, REDG = global reductions).
usage: python scripts/sass_histogram.py > profiles/rNN_sass_opcodes.txt   (runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "physicsml_b200", "csrc", "libphysicsml_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL",
        "REDG", "RED", "ATOMG", "LDGSTS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "MUFU", "HMMA", "IMMA"]

proc = subprocess.Popen(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True)
fam = None
hist = collections.defaultdict(collections.Counter)
nk = collections.Counter()
func = re.compile(r"Function : (\S+)")
ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")
for line in proc.stdout:
    m = func.search(line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() if False else m.group(1)
        # family = the kernel's base name inside namespace pml (mangled: _ZN3pml<len><name>...)
        mm = re.search(r"_ZN3pml(?:4wide)?(\d+)([A-Za-z0-9_]+)", name)
        fam = mm.group(2)[: int(mm.group(1))] if mm else name[:40]
        nk[fam] += 1
        continue
    m = ins.match(line)
    if m and fam:
        op = m.group(1)
        hist[fam][op] += 1
print(f"# SASS opcode counts per kernel family of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; all template instantiations summed)")
print(f"# {'family':34s} {'kernels':>7s} {'instr':>9s}  " + " ".join(f"{k:>7s}" for k in KEYS))
tot = collections.Counter()
for f in sorted(hist, key=lambda f: -sum(hist[f].values())):
    h = hist[f]
    print(f"  {f:34s} {nk[f]:7d} {sum(h.values()):9d}  " + " ".join(f"{h.get(k, 0):7d}" for k in KEYS))
    tot.update(h)
print(f"  {'TOTAL':34s} {sum(nk.values()):7d} {sum(tot.values()):9d}  " + " ".join(f"{tot.get(k, 0):7d}" for k in KEYS))
