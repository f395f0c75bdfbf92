"""Correctness (against the default plan) and timing of an alternate FFT plan selected by an environment variable.
usage: JDSP_FFT_...=1 python tools/fft_plan_check.py <log2 n> [log2 total points]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from jeicyboodsp_b200.binding import Context, Library
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
n = 1 << int(sys.argv[1]); total = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 28)
x = torch.randn(total, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x)
for fwd in (True, False):
    ctx.fft_c2c_f32(x, y, n, total // n, fwd); torch.cuda.synchronize()
    b = min(64, total // n)
    ref = torch.fft.fft(x[: b * n].view(b, n).to(torch.complex128), dim=1) if fwd else torch.fft.ifft(x[: b * n].view(b, n).to(torch.complex128), dim=1) * n
    err = (y[: b * n].view(b, n).to(torch.complex128) - ref).abs().max().item() / ref.abs().max().item()
    last = y[-n:].to(torch.complex128); refl = torch.fft.fft(x[-n:].to(torch.complex128)) if fwd else torch.fft.ifft(x[-n:].to(torch.complex128)) * n
    errl = (last - refl).abs().max().item() / refl.abs().max().item()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.fft_c2c_f32(x, y, n, total // n, fwd); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[3]
    print(f"n={n} fwd={fwd} err {err:.2e} (last transform {errl:.2e})  {ms:.3f} ms  frac {total * 16 / ms / 1e6 / PEAK:.3f}", flush=True)
