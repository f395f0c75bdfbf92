"""Tiny run of every kernel for compute-sanitizer (memcheck / racecheck): a few tiles each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from jeicyboodsp_b200 import synth
from jeicyboodsp_b200.binding import SS, WIENER, Context, Library
L = Library(); ctx = Context(L, 0)
dev = torch.device("cuda")
for preset in ("bench", "ref"):
    for m in (SS, WIENER):
        p = L.denoise_params(preset, m); H = p.hop; nb = 21; S = 3
        x = torch.from_numpy(np.stack([synth.denoise_stream(s, nb * H) for s in range(S)])).to(dev)
        out = torch.empty((S, (nb - 2) * H), dtype=torch.int16, device=dev)
        st = ctx.denoise_state(p, S); st.run(x, nb * H, nb, out, (nb - 2) * H); ctx.sync(); st.close()
p = L.fastconv_params("bench"); B = p.block; nb = 9; S = 4
x = torch.from_numpy(np.stack([synth.fastconv_source(s, nb * B) for s in range(S)])).to(dev)
h = np.stack([np.concatenate([synth.hrir_pair(s), np.zeros((2, 1))], axis=1) for s in range(S)])
st = ctx.fastconv_state(p, S, h); out = torch.empty((S, 2, nb * B), dtype=torch.int16, device=dev)
st.run(x, nb * B, nb, out, nb * B); st.reset(); out2 = torch.empty((2, 2, nb * B), dtype=torch.int16, device=dev)
st.run(x, nb * B, nb, out2, nb * B, sources_per_scene=2); ctx.sync()
p = L.mfcc_params("bench"); plan = ctx.mfcc_plan(p); n = 8000
x = torch.from_numpy(np.stack([synth.mfcc_utterance(u, n) for u in range(3)])).to(dev)
feat = torch.empty((3, plan.n_frames(n), 13), dtype=torch.float32, device=dev); plan.run(x, n, 3, n, feat, plan.n_frames(n) * 13); ctx.sync()
for n in (64, 256, 1024, 8192, 16384, 65536):
    z = torch.randn(3 * n, dtype=torch.complex64, device=dev); y = torch.empty_like(z); ctx.fft_c2c_f32(z, y, n, 3, True); ctx.sync()
sig = torch.from_numpy(synth.roundtrip_signal(512 * 9)).to(dev); o = torch.empty_like(sig)
ctx.roundtrip_dev(sig, 512 * 9, o, 512 * 9, None, 0, 512, 1, 9); ctx.sync()
print("sanitize_small done")
