"""Profiler-friendly batched round trip (N = 1024, 2048 replicas of the 10 s signal): two launches, the second one is the one to capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from jeicyboodsp_b200 import synth
from jeicyboodsp_b200.binding import Context, Library
L = Library(); ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
n_fft, reps = 1024, 2048
sig = synth.roundtrip_signal(160_000); nb = -(-len(sig) // n_fft)
pad = np.zeros(nb * n_fft, np.int16); pad[: len(sig)] = sig
x = torch.from_numpy(pad).cuda().unsqueeze(0).repeat(reps, 1).contiguous(); y = torch.empty_like(x)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.roundtrip_dev(x, nb * n_fft, y, nb * n_fft, None, 0, n_fft, reps, nb); e1.record(); torch.cuda.synchronize()
print(f"roundtrip {n_fft}: {e0.elapsed_time(e1):.3f} ms for {reps} x {nb * n_fft} samples")
