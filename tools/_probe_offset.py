"""FFT sweep sizes >= 4096 against the distance between the input and the output buffer (one allocation, output `pad` bytes past
input + 4 GiB): the L2-slice / HBM-channel hash makes some distances slower."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, json, time
from jeicyboodsp_b200.binding import Context, Library
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
total = 1 << 29
IDLE = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
maxpad = (512 << 20) // 8
xy = torch.empty(2 * total + maxpad, dtype=torch.complex64, device="cuda")
x = xy[:total]; torch.view_as_real(x).uniform_(-1, 1)
sizes = (4096, 8192, 16384, 32768, 65536)
print("pad bytes".rjust(12), " ".join(f"{n:>7d}" for n in sizes), flush=True)
for pad_b in (0, 1 << 20, 0, 1 << 20, 0, 4096, 0):
    y = xy[total + pad_b // 8: 2 * total + pad_b // 8]
    fr = []
    for n in sizes:
        if IDLE: torch.cuda.synchronize(); time.sleep(IDLE)
        for _ in range(3): ctx.fft_c2c_f32(x, y, n, total // n, True)
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.fft_c2c_f32(x, y, n, total // n, True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        fr.append(total * 16 / sorted(ts)[3] / 1e6 / PEAK)
    print(f"{pad_b:12d} " + " ".join(f"{f:7.3f}" for f in fr), flush=True)
