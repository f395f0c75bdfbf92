"""Measure the other BASELINE.json configs (1, 3, 4, 5) and the pitch path on one GPU: CUDA-event timing, algorithmic bytes vs the
measured HBM peak, and a parity spot check against the oracle for each.  Writes one JSON object per config.

    python tools/bench_extras.py [--out gpurun_out/extras.json] [--quick]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from jeicyboodsp_b200 import synth  # noqa: E402
from jeicyboodsp_b200.binding import Context, Library  # noqa: E402
from oracle.oracle import MfccParams as OMP  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/extras.json")
ap.add_argument("--quick", action="store_true")
ap.add_argument("--only", default="")
args = ap.parse_args()

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda")
L = Library()
ctx = Context(L, 0, stream=torch.cuda.current_stream().cuda_stream)
o = Oracle()
results = []


def timed(fn, warm=3, reps=7):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def emit(d):
    results.append(d)
    print(json.dumps(d), flush=True)


def want(name):
    return not args.only or name in args.only.split(",")


# ---- config 5: batched c2c size sweep, 2^29 points (4 GiB in + 4 GiB out) -----------------------------------------
if want("sweep"):
    total = (1 << 29) if not args.quick else (1 << 26)
    # one allocation, input in the first half and output in the second: the distance between the two streams is then
    # exactly the buffer size (measured: other placements cost the N >= 4096 kernels 3-8 %, tools/_probe_offset.py)
    xy = torch.empty(2 * total, dtype=torch.complex64, device=dev)
    x, y = xy[:total], xy[total:]
    xr = torch.view_as_real(x)
    g = torch.Generator(device=dev); g.manual_seed(5)
    xr.uniform_(-1, 1, generator=g)
    for lg in range(8, 17):
        n = 1 << lg
        batch = total // n
        row = {"config": "fft_sweep", "n": n, "batch": batch}
        for fwd, nm in ((True, "fwd"), (False, "inv")):
            med, best = timed(lambda: ctx.fft_c2c_f32(x, y, n, batch, fwd))
            row[f"{nm}_ms"] = med
            row[f"{nm}_gbs"] = total * 16 / med / 1e6
            row[f"{nm}_frac_hbm"] = total * 16 / med / 1e6 / PEAK
            row[f"{nm}_gpoints_s"] = total / med / 1e6
        # parity on 4 random rows: FFTProcess restatement for N <= 2^15, numpy beyond (the reference breaks at 2^16)
        ctx.fft_c2c_f32(x, y, n, batch, True)
        torch.cuda.synchronize()
        rows = np.random.default_rng(n).integers(0, batch, 4)
        worst = 0.0
        for r in rows:
            zi = x[r * n:(r + 1) * n].cpu().numpy().astype(np.complex128)
            ref = o.fftprocess(zi, True) if n <= 32768 else np.fft.fft(zi)
            got = y[r * n:(r + 1) * n].cpu().numpy()
            worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
        row["parity_max_rel"] = worst
        emit(row)
    del x, y, xr, xy
    torch.cuda.empty_cache()

# ---- config 1: FFT -> IFFT round trip ------------------------------------------------------------------------------
if want("roundtrip"):
    sig = synth.roundtrip_signal(160_000)
    for n_fft in (512, 1024):
        nb = -(-len(sig) // n_fft)
        pad = np.zeros(nb * n_fft, np.int16); pad[: len(sig)] = sig
        reps = 16384 if not args.quick else 512
        one = torch.from_numpy(pad).to(dev)
        x = one.unsqueeze(0).repeat(reps, 1).contiguous()
        y = torch.empty_like(x)
        f32 = torch.empty((1, nb * n_fft), dtype=torch.float32, device=dev)
        med1, best1 = timed(lambda: ctx.roundtrip_dev(one, nb * n_fft, y, nb * n_fft, None, 0, n_fft, 1, nb))
        medb, bestb = timed(lambda: ctx.roundtrip_dev(x, nb * n_fft, y, nb * n_fft, None, 0, n_fft, reps, nb))
        ctx.roundtrip_dev(one, nb * n_fft, y, nb * n_fft, f32, nb * n_fft, n_fft, 1, nb)
        torch.cuda.synchronize()
        ref_i16, ref_f64 = o.roundtrip(pad, n_fft)
        emit({"config": "roundtrip", "n_fft": n_fft, "single_stream_us": med1 * 1e3, "single_stream_msamples_s": nb * n_fft / med1 / 1e3,
              "batched_replicas": reps, "batched_ms": medb, "batched_msamples_s": reps * nb * n_fft / medb / 1e3,
              "batched_gbs": reps * nb * n_fft * 4 / medb / 1e6, "batched_frac_hbm": reps * nb * n_fft * 4 / medb / 1e6 / PEAK,
              "parity_precast_max_rel_peak": float(np.abs(f32[0].cpu().numpy() - ref_f64).max() / np.abs(ref_f64).max()),
              "parity_i16_max_lsb": int(np.abs(y[0].cpu().numpy().astype(int) - ref_i16.astype(int)).max())})
        del x, y
    torch.cuda.empty_cache()

# ---- config 3: fast convolution, 512-tap HRIR pairs, 48 kHz ---------------------------------------------------------
if want("fastconv"):
    S = 16384 if not args.quick else 1024
    n = 480_000 if not args.quick else 48_000
    p = L.fastconv_params("bench")
    B = p.block
    nb = n // B
    n = nb * B
    g = torch.Generator(device=dev); g.manual_seed(3)
    x = torch.empty((S, n), dtype=torch.int16, device=dev)
    t = torch.arange(n, device=dev, dtype=torch.float32) / 48000.0
    for s0 in range(0, S, 512):
        s1 = min(S, s0 + 512)
        f = (300.0 + 7.0 * (torch.arange(s0, s1, device=dev) % 64)).to(torch.float32)[:, None]
        v = 2000.0 * torch.randn((s1 - s0, n), generator=g, device=dev) + 2500.0 * torch.sin(2 * np.pi * f * t[None, :])
        x[s0:s1] = torch.clamp(torch.round(v), -32768, 32767).to(torch.int16)
        del v
    rng = np.random.default_rng(3)
    k = np.arange(512)
    h = rng.normal(0, 0.35, (S, 2, 512)) * np.exp(-k / 60.0)[None, None, :]
    h[:, :, :9] = 0.0
    h[:, :, 8] = 1.0
    h *= np.minimum(1.0, 3.0 / np.abs(h).sum(axis=2, keepdims=True))
    taps = np.concatenate([h, np.zeros((S, 2, 1))], axis=2)
    st = ctx.fastconv_state(p, S, taps)
    out = torch.empty((S, 2, (nb - 1) * B), dtype=torch.int16, device=dev)

    def run():
        st.reset()
        st.run(x, n, nb, out, (nb - 1) * B)
    med, best = timed(run, warm=2, reps=5)
    alg = S * n * 2 + S * 2 * (nb - 1) * B * 2
    run(); torch.cuda.synchronize()
    worst, flips, tot = 0, 0, 0
    for s in (0, S // 2, S - 1):
        xs = x[s].cpu().numpy()
        got = out[s].cpu().numpy()
        for ear in range(2):
            ref, _ = o.fastconv(xs, h[s, ear], B, 1, 1024)
            d = np.abs(got[ear].astype(int) - ref.astype(int))
            worst, flips, tot = max(worst, int(d.max())), flips + int((d > 0).sum()), tot + d.size
    emit({"config": "fastconv", "sources": S, "samples_per_source": n, "ms": med, "input_msamples_s": S * n / med / 1e3,
          "algorithmic_bytes": alg, "gbs": alg / med / 1e6, "frac_hbm": alg / med / 1e6 / PEAK,
          "parity_i16_max_lsb": worst, "parity_flip_fraction": flips / tot})
    st.close()
    del x, out
    torch.cuda.empty_cache()

# ---- config 4: MFCC, 400/160/512/26/13 -------------------------------------------------------------------------------
if want("mfcc"):
    U = 36000 if not args.quick else 2048
    n = 160_000
    p = L.mfcc_params("bench")
    plan = ctx.mfcc_plan(p)
    nf = plan.n_frames(n)
    x = synth.denoise_streams_torch(U, n, dev, sigma=25.0, seed=4)
    feat = torch.empty((U, nf, 13), dtype=torch.float32, device=dev)
    med, best = timed(lambda: plan.run(x, n, U, n, feat, nf * 13), warm=2, reps=5)
    alg = U * n * 2 + U * nf * 13 * 4
    torch.cuda.synchronize()
    worst = 0.0
    for u in (0, U // 2, U - 1):
        ref = o.mfcc_frames(x[u].cpu().numpy(), OMP.preset("bench"))
        worst = max(worst, float(np.abs(feat[u].cpu().numpy() - ref).max() / np.abs(ref).max()))
    emit({"config": "mfcc", "utterances": U, "samples_per_utt": n, "frames": U * nf, "ms": med, "msamples_s": U * n / med / 1e3,
          "frames_per_s": U * nf / med * 1e3, "algorithmic_bytes": alg, "gbs": alg / med / 1e6, "frac_hbm": alg / med / 1e6 / PEAK,
          "parity_max_rel_peak": worst})
    plan.close()

# ---- pitch by FFT autocorrelation, 1024-pt frames, hop 512 ------------------------------------------
if want("pitch"):
    S = 4096 if not args.quick else 512
    n = 960_000 if not args.quick else 96_000
    p = L.pitch_params("ref")
    H = p.block
    nb = n // H
    n = nb * H
    x = synth.denoise_streams_torch(S, n, dev)
    arg = torch.empty((S, nb), dtype=torch.int32, device=dev)
    rmax = torch.empty((S, nb), dtype=torch.float64, device=dev)
    st = ctx.pitch_state(p, S)

    def run_pitch():
        st.reset()
        st.run(x, n, nb, arg, rmax)
    med, best = timed(run_pitch, warm=2, reps=5)
    alg = S * n * 2 + S * nb * 12
    torch.cuda.synchronize()
    same = True
    for s_ in (0, S // 2, S - 1):
        ea, em = o.pitch(x[s_].cpu().numpy(), exact=True)
        same = same and bool(np.array_equal(arg[s_].cpu().numpy(), ea)) and bool(np.array_equal(rmax[s_].cpu().numpy(), em))
    emit({"config": "pitch", "streams": S, "samples_per_stream": n, "frames": S * nb, "ms": med, "msamples_s": S * n / med / 1e3,
          "frames_per_s": S * nb / med * 1e3, "algorithmic_bytes": alg, "gbs": alg / med / 1e6, "frac_hbm": alg / med / 1e6 / PEAK,
          "parity_arg_and_rmax_bit_exact": same})
    st.close()
    del x

# ---- two-microphone MVDR, 1024-pt frames, hop 512 --------------------------------------------------
if want("mvdr"):
    S = 4096 if not args.quick else 512
    n = 960_000 if not args.quick else 96_000
    p = L.mvdr_params("ref")
    B = p.block
    nb = n // B
    n = nb * B
    xl = synth.denoise_streams_torch(S, n, dev)
    g = torch.Generator(device=dev); g.manual_seed(6)
    xr = torch.empty_like(xl)
    for s0 in range(0, S, 256):    # right microphone: the scene 3 samples later at 0.8, plus its own sensor noise
        late = torch.roll(xl[s0:s0 + 256].to(torch.float32), 3, dims=1)
        late[:, :3] = 0
        late = 0.8 * late + 35.0 * torch.randn(late.shape, generator=g, device=dev)
        xr[s0:s0 + 256] = torch.clamp(torch.round(late), -32768, 32767).to(torch.int16)
        del late
    out = torch.empty((S, (nb - 1) * B), dtype=torch.int16, device=dev)
    # delay 0 = the program's configuration (single-pass time-domain kernel); a steered beam takes the transform kernels
    for name, dtime in (("mvdr", 0.0), ("mvdr_steered", 2.5e-4)):
        p.dtime = dtime
        st = ctx.mvdr_state(p, S)

        def run_mvdr():
            st.reset()
            st.run(xl, xr, n, nb, out, (nb - 1) * B)
        med, best = timed(run_mvdr, warm=2, reps=5)
        alg = 2 * S * n * 2 + S * (nb - 1) * B * 2
        torch.cuda.synchronize()
        worst, flips, tot = 0, 0, 0
        for s_ in (0, S // 2, S - 1):
            ref = o.mvdr(xl[s_].cpu().numpy(), xr[s_].cpu().numpy(), dtime)[0]
            d = np.abs(out[s_].cpu().numpy().astype(int) - ref.astype(int))
            worst, flips, tot = max(worst, int(d.max())), flips + int((d > 0).sum()), tot + d.size
        emit({"config": name, "steering_delay_s": dtime, "streams": S, "samples_per_microphone": n, "frames": S * nb, "ms": med,
              "msamples_s": S * n / med / 1e3, "frames_per_s": S * nb / med * 1e3, "algorithmic_bytes": alg, "gbs": alg / med / 1e6,
              "frac_hbm": alg / med / 1e6 / PEAK, "parity_i16_max_lsb": worst, "parity_flip_fraction": flips / tot})
        st.close()
    del xl, xr, out

os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w") as f:
    json.dump({"peak_hbm_gbs": PEAK, "results": results}, f, indent=1)
