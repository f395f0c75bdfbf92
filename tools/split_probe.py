"""N = 32768: four-step plan vs the radix-2 split over the on-chip 16384 kernel (the default; JDSP_FFT_NO_SPLIT=1 selects the four-step). 4 GiB in + 4 GiB out."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from jeicyboodsp_b200.binding import Context, Library
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6454.6
L = Library(sys.argv[1] if len(sys.argv) > 1 else None); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
total = 1 << 29
xy = torch.empty(2 * total, dtype=torch.complex64, device="cuda"); x, y = xy[:total], xy[total:]
torch.view_as_real(x).uniform_(-1, 1)
n = 32768; batch = total // n
for plan in ("fourstep", "split", "fourstep", "split"):
    if plan == "fourstep": os.environ["JDSP_FFT_NO_SPLIT"] = "1"
    else: os.environ.pop("JDSP_FFT_NO_SPLIT", None)
    for fwd in (True, False):
        ts = []
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.fft_c2c_f32(x, y, n, batch, fwd); e1.record(); torch.cuda.synchronize()
            if it >= 3: ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        r = np.random.default_rng(1).integers(0, batch, 2)
        err = 0.0
        for q in r:
            zi = x[q * n:(q + 1) * n].cpu().numpy().astype(np.complex128)
            ref = np.fft.fft(zi) if fwd else np.fft.ifft(zi) * n
            err = max(err, float(np.abs(y[q * n:(q + 1) * n].cpu().numpy() - ref).max() / np.abs(ref).max()))
        print(f"{plan:9s} fwd={int(fwd)}  {ms:.3f} ms  {total * 16 / ms / 1e6:.0f} GB/s  frac {total * 16 / ms / 1e6 / PEAK:.3f}  max rel err {err:.2e}", flush=True)
