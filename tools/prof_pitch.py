"""Timing run of the pitch kernel (PitchEstimation_method1 framing) on device-resident streams."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from jeicyboodsp_b200 import synth  # noqa: E402
from jeicyboodsp_b200.binding import Context, Library  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=4096)
ap.add_argument("--seconds", type=float, default=20.0)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
L = Library()
ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
p = L.pitch_params("ref")
H = p.block
nb = int(a.seconds * 16000) // H
n = nb * H
x = synth.denoise_streams_torch(a.streams, n, torch.device("cuda"))
arg = torch.empty((a.streams, nb), dtype=torch.int32, device="cuda")
rmax = torch.empty((a.streams, nb), dtype=torch.float64, device="cuda")
st = ctx.pitch_state(p, a.streams)
for it in range(a.iters):
    st.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.run(x, n, nb, arg, rmax)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"iter {it}: {ms:.3f} ms  {a.streams * n / ms / 1e3:.1f} Msamples/s  {a.streams * nb / ms / 1e3:.2f} Mframes/s  "
          f"{a.streams * (n * 2 + nb * 12) / ms / 1e6:.1f} GB/s algorithmic")
print("arg histogram head:", torch.bincount(arg.flatten().clamp(0, 511))[100:110].tolist())
