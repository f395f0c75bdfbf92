import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from jeicyboodsp_b200 import synth
from jeicyboodsp_b200.binding import Context, Library
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
dev = torch.device("cuda")
U, n = 4500, 160000
plan = ctx.mfcc_plan(L.mfcc_params("bench")); nf = plan.n_frames(n)
x = synth.denoise_streams_torch(U, n, dev, sigma=25.0, seed=4)
def timed(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
feat = torch.empty((U, nf * 13), dtype=torch.float32, device=dev)
print("plain", timed(lambda: plan.run(x, n, U, n, feat, nf * 13)))
for nd in (1, 2, 4, 8):
    mats = [torch.empty((U, nf * 13), dtype=torch.float32, device=dev) for _ in range(nd)]
    print("scatter to", nd, "local matrices", timed(lambda: plan.run_scatter(x, n, U, n, mats, nf * 13)))
    del mats
