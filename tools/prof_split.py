"""Small instance of the N = 32768 transform for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from jeicyboodsp_b200.binding import Context, Library
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
n, batch = 32768, 4096
x = torch.randn(n * batch, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x)
for _ in range(3): ctx.fft_c2c_f32(x, y, n, batch, True)
torch.cuda.synchronize()
