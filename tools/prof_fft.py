import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from jeicyboodsp_b200.binding import Context, Library
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
total = 1 << 27
x = torch.randn(total, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.fft_c2c_f32(x, y, n, total // n, True); e1.record(); torch.cuda.synchronize()
    print(n, e0.elapsed_time(e1), "ms", total * 16 / e0.elapsed_time(e1) / 1e6, "GB/s")
