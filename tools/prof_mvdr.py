"""Timing run of the MVDR kernels (BeamForming_MVDR_ver1 framing) on device-resident microphone pairs."""
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
ap.add_argument("--dtime", type=float, default=0.0)
ap.add_argument("--lib", default=None)
a = ap.parse_args()
L = Library(a.lib)
ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
p = L.mvdr_params("ref")
p.dtime = a.dtime
B = p.block
nb = int(a.seconds * 16000) // B
n = nb * B
xl = synth.denoise_streams_torch(a.streams, n, torch.device("cuda"))
xr = torch.roll(xl, 3, dims=1).contiguous()
xr = (0.8 * xr.to(torch.float32) + 35.0 * torch.randn(xr.shape, device="cuda")).round().clamp(-32768, 32767).to(torch.int16)
out = torch.empty((a.streams, (nb - 1) * B), dtype=torch.int16, device="cuda")
st = ctx.mvdr_state(p, a.streams)
for it in range(a.iters):
    st.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.run(xl, xr, n, nb, out, (nb - 1) * B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"iter {it}: {ms:.3f} ms  {a.streams * n / ms / 1e3:.1f} Msamples/s per microphone  {a.streams * nb / ms / 1e3:.2f} Mframes/s  "
          f"{a.streams * n * 6 / ms / 1e6:.1f} GB/s algorithmic")
print("spatial matrix of stream 0:", st.spatial_corr()[0].tolist(), " output rms:", out[:8].float().pow(2).mean().sqrt().item())
