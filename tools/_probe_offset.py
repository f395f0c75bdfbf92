import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, json
from jeicyboodsp_b200.binding import Context, Library
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
total = 1 << 29
for mode in ("empty+uniform", "randn"):
    if mode == "randn":
        x = torch.randn(total, dtype=torch.complex64, device="cuda")
    else:
        x = torch.empty(total, dtype=torch.complex64, device="cuda"); torch.view_as_real(x).uniform_(-1, 1)
    for pad in (0, 4096 // 8, (1 << 20) // 8 + 512, (37 << 20) // 8):
        ybuf = torch.empty(total + pad, dtype=torch.complex64, device="cuda"); y = ybuf[pad:]
        print(mode, "pad elems", pad, "x", hex(x.data_ptr()), "y", hex(y.data_ptr()), "diff MiB", (y.data_ptr() - x.data_ptr()) / 2**20, flush=True)
        for n in (4096, 8192, 16384):
            for _ in range(3): ctx.fft_c2c_f32(x, y, n, total // n, True)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ctx.fft_c2c_f32(x, y, n, total // n, True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[2]
            print(f"   n={n:6d} {ms:7.3f} ms frac {total*16/ms/1e6/PEAK:.3f}", flush=True)
        del ybuf, y
    del x
    torch.cuda.empty_cache()
