"""Aggregate an ncu SASS-page CSV per source line, using nvdisasm line info of the built cubin.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel mangled substring> [top]"""
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "jeicyboodsp_b200", "libjdsp.so")], cwd=d, check=True, stdout=subprocess.DEVNULL)
    dis = ""
    for cub in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):   # one cubin per translation unit
        dis += subprocess.run(["nvdisasm", "--print-line-info", "--print-code", os.path.join(d, cub)], capture_output=True, text=True).stdout
# address -> line
sec, line, amap = None, None, {}
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        sec = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and sec and ksub in sec:
        amap[int(m.group(1), 16)] = (line, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur, hdr, per_line, per_op, tot_i, tot_s, tot_w = None, None, {}, {}, 0.0, 0.0, 0.0
first_addr = None
nk = 0
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = r[1]; nk += 1; first_addr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr) or cur is None:
        continue
    if nk != which:
        continue
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    if first_addr is None: first_addr = a
    a = a - first_addr
    ie = float(r[hdr.index("Instructions Executed")] or 0)
    sm = float(r[hdr.index("# Samples")] or 0)
    wf = float(r[hdr.index("L1 Wavefronts Shared")] or 0) if "L1 Wavefronts Shared" in hdr else 0.0
    tot_i += ie; tot_s += sm; tot_w += wf
    base = min(amap) if amap else 0
    key, op = amap.get(a, amap.get(a - 0, ((None, 0), r[1])))
    per_line.setdefault(key, [0.0, 0.0, 0.0]); per_line[key][0] += ie; per_line[key][1] += sm; per_line[key][2] += wf
    opn = r[1].split()[0] if not r[1].startswith("@") else r[1].split()[1]
    opn = opn.split(".")[0] + ("." + opn.split(".")[1] if opn.startswith(("LDS", "STS", "LDG", "STG")) and "." in opn else "")
    per_op.setdefault(opn, [0.0, 0.0, 0.0]); per_op[opn][0] += ie; per_op[opn][1] += sm; per_op[opn][2] += wf
print(f"total warp-instr {tot_i:.0f}, samples {tot_s:.0f}, shared wavefronts {tot_w:.0f}")
src_cache = {}
def src(key):
    if not key or not key[0]: return "?"
    f, l = key
    if f not in src_cache:
        p = os.path.join(ROOT, "jeicyboodsp_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    t = src_cache[f]
    return f"{f}:{l}  " + (t[l - 1].strip()[:100] if 0 < l <= len(t) else "")
print("--- by source line")
for key, (ie, sm, wf) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ie / tot_i * 100:5.1f}% inst {sm / max(tot_s, 1) * 100:5.1f}% samples {wf / max(tot_w, 1) * 100:5.1f}% smem-wf | {src(key)}")
print("--- by opcode")
for op, (ie, sm, wf) in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{ie / tot_i * 100:5.1f}% inst {sm / max(tot_s, 1) * 100:5.1f}% samples {wf / max(tot_w, 1) * 100:5.1f}% smem-wf | {op}")
