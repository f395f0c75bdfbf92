"""Average active threads per executed warp instruction, per kernel of an ncu report (source page): a warp that runs split
(e.g. lane 0 apart from lanes 1..31 after a single-lane branch at a loop boundary) shows up as a ratio well below 32.
usage: python tools/ncu_divergence.py <report.ncu-rep> [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur, hdr, acc = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = r[1]; acc.setdefault(cur, []); continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur is None or hdr is None or len(r) < len(hdr):
        continue
    try:
        wi = float(r[hdr.index("Instructions Executed")]); ti = float(r[hdr.index("Thread Instructions Executed")])
    except ValueError:
        continue
    acc[cur].append((wi, ti, r[hdr.index("Source")].strip()))
for k, v in acc.items():
    W = sum(a for a, _, _ in v); T = sum(b for _, b, _ in v)
    print(f"{k[:90]}\n   warp instructions {W:.0f}, threads per instruction {T / max(W, 1):.2f}")
    low = sorted((x for x in v if x[0] > 0.002 * W and x[1] / x[0] < 24), key=lambda x: -x[0])[:top]
    for wi, ti, src in low:
        print(f"      {wi / W * 100:5.2f}% of instructions at {ti / wi:5.1f} threads: {src[:80]}")
