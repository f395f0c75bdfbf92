import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from jeicyboodsp_b200 import synth
from jeicyboodsp_b200.binding import SS, Context, Library
L = Library(); ctx = Context(L, 0)
S, n = 2048, 960000
p = L.denoise_params("bench", SS)
nb = n // 256; n_out = (nb - 2) * 256
x = synth.denoise_streams_torch(256, n, torch.device("cuda")).cpu()
h_in = torch.empty((S, n), dtype=torch.int16).pin_memory(); h_out = torch.empty((S, n_out), dtype=torch.int16).pin_memory()
for i in range(S // 256): h_in[i * 256:(i + 1) * 256] = x
for it in range(3):
    t0 = time.perf_counter(); ctx.denoise_host_raw(p, h_in, n, S, n, h_out, n_out); dt = time.perf_counter() - t0
    print(f"e2e pass: {dt*1e3:.1f} ms  {S*n/dt/1e6:.0f} Msamples/s  {(S*n*2+S*n_out*2)/dt/1e9:.1f} GB/s over PCIe")
