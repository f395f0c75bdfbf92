"""Profiler-friendly small instances of the other kernels: mfcc, fastconv, fft sweep sizes."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from jeicyboodsp_b200 import synth
from jeicyboodsp_b200.binding import Context, Library
ap = argparse.ArgumentParser()
ap.add_argument("--which", default="mfcc,fastconv,fft")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda")
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
def t(fn, label, units):
    for i in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1):.3f} ms  {units / e0.elapsed_time(e1) / 1e3:.1f} Munits/s", flush=True)
if "mfcc" in a.which:
    U, n = 2960, 160000
    p = L.mfcc_params("bench"); plan = ctx.mfcc_plan(p); nf = plan.n_frames(n)
    x = synth.denoise_streams_torch(U, n, dev, sigma=25.0, seed=4)
    feat = torch.empty((U, nf, 13), dtype=torch.float32, device=dev)
    t(lambda: plan.run(x, n, U, n, feat, nf * 13), "mfcc", U * n)
if "fastconv" in a.which:
    S, nb = 1184, 200
    p = L.fastconv_params("bench"); B = p.block; n = nb * B
    x = (2000 * torch.randn((S, n), device=dev)).round().clamp(-32768, 32767).to(torch.int16)
    taps = np.zeros((S, 2, 513)); taps[:, :, 8] = 1.0; taps[:, :, 9:200] = np.random.default_rng(0).normal(0, 0.02, (S, 2, 191))
    st = ctx.fastconv_state(p, S, taps)
    out = torch.empty((S, 2, (nb - 1) * B), dtype=torch.int16, device=dev)
    def run():
        st.reset(); st.run(x, n, nb, out, (nb - 1) * B)
    t(run, "fastconv", S * n)
if "fft" in a.which:
    total = 1 << 26
    x = torch.randn(total, dtype=torch.complex64, device=dev); y = torch.empty_like(x)
    for n in (1024, 8192, 16384, 65536):
        t(lambda: ctx.fft_c2c_f32(x, y, n, total // n, True), f"fft{n}", total)
