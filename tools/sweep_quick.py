import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, json
from jeicyboodsp_b200.binding import Context, Library
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
total = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
x = torch.randn(total, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x)
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 8
for lg in range(lo, 17):
    n = 1 << lg
    for _ in range(3): ctx.fft_c2c_f32(x, y, n, total // n, True)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.fft_c2c_f32(x, y, n, total // n, True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"n={n:6d} {ms:7.3f} ms {total*16/ms/1e6:8.1f} GB/s frac {total*16/ms/1e6/PEAK:.3f}", flush=True)
