"""Per-kernel SASS opcode counts of libjdsp.so (cuobjdump -sass): the Blackwell-specific instructions the kernels rely on --
UBLKCP (1-D bulk TMA copy), SYNCS (mbarrier), LDGSTS (cp.async), FFMA2 / FADD2 / FMUL2 (packed f32x2), REDUX, SHFL, MUFU,
UCGABAR / cluster barriers and st.shared::cluster (distributed shared memory).  usage: python tools/sass_counts.py > profiles/.../sass_opcode_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "jeicyboodsp_b200", "libjdsp.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DMUL", "REDUX", "SHFL", "MUFU", "LDS", "STS", "LDG", "STG",
         "UCGABAR_ARV", "UCGABAR_WAIT", "ST.E", "LDL", "STL", "BAR"]
kern, counts, arch = None, collections.OrderedDict(), set()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("jdsp::", "")
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        if op in WATCH:
            counts[kern][op] += 1
        if op == "ST" and m.group(2) and "SHARED" in (m.group(2) or ""):
            counts[kern]["ST.SHARED.CLUSTER"] += 1
print(f"libjdsp.so: {len(counts)} kernels, arch {sorted(arch)}")
cols = ["total", "FFMA2", "FADD2", "FMUL2", "FFMA", "UBLKCP", "SYNCS", "LDGSTS", "SHFL", "REDUX", "MUFU", "LDS", "STS", "LDL", "STL", "UCGABAR_ARV", "UCGABAR_WAIT", "BAR"]
print(f"{'kernel':78s} " + " ".join(f"{c:>7s}" for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    print(f"{k[:78]:78s} " + " ".join(f"{c.get(col, 0):7d}" for col in cols))
    tot.update(c)
print(f"{'ALL':78s} " + " ".join(f"{tot.get(col, 0):7d}" for col in cols))
