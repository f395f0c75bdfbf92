"""Small, profiler-friendly run of the fused denoise kernel (same kernel and preset as bench.py, fewer streams)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from jeicyboodsp_b200 import synth  # noqa: E402
from jeicyboodsp_b200.binding import SS, WIENER, Context, Library  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1184)
ap.add_argument("--seconds", type=float, default=8.0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--preset", default="bench")
a = ap.parse_args()
L = Library()
ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
p = {m: L.denoise_params(a.preset, m) for m in (SS, WIENER)}
H = p[SS].hop
nb = int(a.seconds * 16000) // H
n = nb * H
x = synth.denoise_streams_torch(a.streams, n, torch.device("cuda"))
out = torch.empty((a.streams, (nb - 2) * H), dtype=torch.int16, device="cuda")
st = {m: ctx.denoise_state(p[m], a.streams) for m in (SS, WIENER)}
for it in range(a.iters):
    for m in (SS, WIENER):
        st[m].reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st[m].run(x, n, nb, out, (nb - 2) * H)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"iter {it} mode {m}: {ms:.3f} ms  {a.streams * n / ms / 1e3:.1f} Msamples/s  "
              f"{a.streams * (n + (nb - 2) * H) * 2 / ms / 1e6:.1f} GB/s algorithmic")
