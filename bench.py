#!/usr/bin/env python
"""bench.py -- the frame-wise spectral hot path on B200 (BASELINE.json metric), every BASELINE config in one line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--only denoise,fastconv,...]

Headline (`metric` / `value` / `config` / `roofline` / `e2e` / `cpu_baseline`): BASELINE.json configs[1] -- spectral
subtraction + Wiener denoise, 512-pt Hann STFT, 50 % overlap, 16 kHz synthetic speech + AWGN, 4096 streams x 60 s PER GPU
(weak scaling: streams are independent, no collective on the data path).  One step = one fused spectral-subtraction pass
plus one fused Wiener pass over every stream (two kernel launches).  `value` counts samples through a denoiser (streams x
samples x 2 passes) per second, inputs resident in HBM; `e2e` is the same through the host-buffer C-ABI call
(jdsp_denoise_i16: pinned host -> device -> pinned host, copies inside the timed region).

`configs`: the other four BASELINE.json configs, each with its own CUDA-event time, HBM and fp32 roofline fractions, in-run
parity against the oracle, an end-to-end number through the host-buffer C-ABI call and the matching reference program timed
on the host cores.  Their TOTAL workload is fixed (BASELINE's sizes) and sharded 1/N per rank under --gpus N ("strong");
MFCC additionally reports the NCCL all-gather of the per-GPU feature blocks, serial and overlapped with compute.

--impl reference times the reference's own CPU programs (oracle/_ref, compiled from the unmodified sources; FFTW calls
served by oracle/fftw_shim) on the host cores, one process per core, on a bounded sample of the headline workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 16_000
N_FFT, HOP = 512, 256
METRIC = "Msamples/s through fused STFT denoise (spectral subtraction + Wiener)"
# CUDA-core fp32 peak: measured FFMA issue rate 3.88 warp-instructions / clk / SM (profiles/microbench/mb.txt) x 32 lanes x
# 2 flop x 148 SMs x 1.965 GHz.  FFMA2 (packed f32x2) issues at half that rate with twice the flops: same peak.
FP32_PEAK_TFLOPS = 72.6
# flop per unit of the 5 N log2 N model with real-input packing: what the fp32 roof is computed from
FLOP_PER_SAMPLE = {"denoise": 108.0, "fastconv": 165.0, "mfcc": 105.0, "roundtrip": 53.0}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _roof(kernel, alg_bytes, ms, flop, note=None):
    """HBM fraction (the judged figure) and the fp32-issue fraction next to it (configs 1-4 are bound by the
    CUDA-core fp32 rate, only the complex64 sweep by HBM)."""
    peak, src = _peaks()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    tf = flop / (ms * 1e-3) / 1e12
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": src,
         "fp32_tflops": tf, "fp32_peak_tflops": FP32_PEAK_TFLOPS, "fp32_frac": tf / FP32_PEAK_TFLOPS, "kernel": kernel,
         "algorithmic_bytes_per_launch": int(alg_bytes), "avg_launch_ms": ms}
    if note:
        r["note"] = note
    return r


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own programs on the host cores
def _cpu_workers():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _run_procs(cmds):
    """One process alone, then all of them at once: (seconds single, seconds all)."""
    t0 = time.perf_counter()
    subprocess.run(cmds[0], check=True, stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t_single = time.perf_counter() - t0
    t0 = time.perf_counter()
    procs = [subprocess.Popen(c, stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
    for p in procs:
        p.wait()
    return t_single, time.perf_counter() - t0


def run_reference_cpu(streams_per_core: int, seconds: float, cores: int | None = None):
    """One process per core, each denoising `streams_per_core` private streams with the SS program and the same
    number with the Wiener program (bench preset).  Returns (Msamples/s aggregate, cores, kind, sample text,
    single-process Msamples/s)."""
    import numpy as np
    from jeicyboodsp_b200 import synth
    from oracle.oracle import REF_DIR, Oracle, RefPrograms

    cores = cores or _cpu_workers()
    n = int(seconds * FS)
    ref = RefPrograms()
    kind = "reference" if ref.available("ss_bench") and ref.available("wiener_bench") else "port"
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        x = synth.denoise_stream(0, n)
        files = []
        for c in range(cores):
            fi = os.path.join(d, f"in_{c}.pcm")
            np.roll(x, 977 * c).tofile(fi)
            files.append(fi)
        if kind == "reference":
            script = ("for i in $(seq %d); do %s/ss_bench $1 $1.ss </dev/null >/dev/null 2>&1; "
                      "%s/wiener_bench $1 $1.wf </dev/null >/dev/null 2>&1; done" % (streams_per_core, REF_DIR, REF_DIR))
            cmds = [["bash", "-c", script, "bash", f] for f in files]
        else:
            Oracle()  # make sure the restatement is built before the workers start
            code = ("import sys,numpy as np; sys.path.insert(0,%r); from oracle.oracle import Oracle, DenoiseParams as P;"
                    "o=Oracle(); x=np.fromfile(sys.argv[1],np.int16);"
                    "[(o.denoise(x,P.preset('bench',0)),o.denoise(x,P.preset('bench',1))) for _ in range(%d)]" % (ROOT, streams_per_core))
            cmds = [[sys.executable, "-c", code, f] for f in files]
        t_single, t_all = _run_procs(cmds)
    samples_per_proc = 2 * streams_per_core * n
    agg = cores * samples_per_proc / t_all / 1e6
    single = samples_per_proc / t_single / 1e6
    what = (f"{cores} processes x {streams_per_core} streams x {seconds:.0f} s x 2 programs (ss_bench, wiener_bench: "
            f"512-pt Hann, hop 256), files in tmpfs, stdout discarded; "
            + ("unmodified reference sources, FFTW calls served by a radix-2 double shim (not FFTW)" if kind == "reference"
               else "oracle C restatement (oracle/_ref absent)"))
    return agg, cores, kind, what, single, t_all


def _cpu_program_baseline(exe: str, make_inputs, argv, samples_per_proc: int, what: str, reps: int = 1, fftw_shim: bool = True):
    """Time one reference console program (oracle/_ref/<exe>) as one process per core on private tmpfs files."""
    from oracle.oracle import REF_DIR, RefPrograms
    if not RefPrograms().available(exe):
        return {"unavailable": f"oracle/_ref/{exe} not built"}
    cores = _cpu_workers()
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        cmds = []
        for c in range(cores):
            files = make_inputs(d, c)
            one = " ".join([os.path.join(REF_DIR, exe)] + argv(files)) + " </dev/null >/dev/null 2>&1"
            cmds.append(["bash", "-c", "; ".join([one] * reps)])
        t_single, t_all = _run_procs(cmds)
    return {"value": cores * samples_per_proc * reps / t_all / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
            "single_process_msamples_s": samples_per_proc * reps / t_single / 1e6,
            "sample": f"{cores} processes x {what}; unmodified reference source"
                      + (", FFTW calls served by a radix-2 double shim (not FFTW)" if fftw_shim else "")}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        agg, cores, kind, what, single, t_all = run_reference_cpu(args.ref_streams_per_core, args.ref_seconds)
        if i >= args.warmup:
            vals.append(agg)
            times.append(t_all)
    v = statistics.median(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "denoise: SS + Wiener, 512-pt Hann STFT, hop 256, 16 kHz speech+AWGN (bounded CPU sample)",
                   "n_fft": N_FFT, "hop": HOP},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": what,
                         "single_process_msamples_s": single},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
class Bench:
    """Shared plumbing of the GPU arm: ranks, timers, pinned buffers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: libjdsp has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from jeicyboodsp_b200.binding import Context, Library
        self.L = Library()
        self.ctx = Context(self.L, self.local, stream=torch.cuda.current_stream().cuda_stream)
        self.ectx = None       # context of the host-buffer forms (own stream + pipe streams)
        self._oracle = None

    @property
    def oracle(self):
        if self._oracle is None:
            from oracle.oracle import Oracle
            self._oracle = Oracle()
        return self._oracle

    def host_ctx(self):
        if self.ectx is None:
            from jeicyboodsp_b200.binding import Context
            self.ectx = Context(self.L, self.local)
        return self.ectx

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ranks(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_ranks_bool(self, ok: bool) -> bool:
        """True only if `ok` holds on every rank."""
        return self.max_ranks(0.0 if ok else 1.0) == 0.0

    def timed(self, fn, warm=3, reps=5, idle_s=0.0):
        """Mean device time of fn (CUDA events on the launching stream), max over ranks; inputs of every config exceed L2.
        idle_s > 0: the device is left idle that long before the warm-up launches (a kernel timed on its own, not inside a long run)."""
        torch = self.torch
        if idle_s > 0:
            torch.cuda.synchronize()
            time.sleep(idle_s)
        for _ in range(warm):
            fn()
        self.barrier()
        evs = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) for a, b in evs]
        return self.max_ranks(statistics.mean(ts)), min(ts)

    def wall(self, fn, reps=2):
        """Host wall clock around blocking host-buffer calls, max over ranks (seconds per call)."""
        fn()   # warm-up: workspace allocation, stream creation
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        self.torch.cuda.synchronize()
        return self.max_ranks((time.perf_counter() - t0) / reps)

    def free(self):
        self.torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------
def cfg_fastconv(B: Bench):
    """BASELINE configs[2]: 48 kHz mono sources x 512-tap HRIR pairs, overlap-save, 16384 sources -> binaural (mode A)."""
    import numpy as np
    torch, dev, L, args = B.torch, B.dev, B.L, B.args
    from jeicyboodsp_b200.sharding import shard_range
    S_total, n = (16384, 480_000) if not args.quick else (1024, 48_000)
    s_lo, s_hi = shard_range(S_total, B.rank, B.world)
    S = s_hi - s_lo
    p = L.fastconv_params("bench")
    Bk = p.block
    nb = n // Bk
    n = nb * Bk
    n_out = (nb - 1) * Bk
    g = torch.Generator(device=dev); g.manual_seed(3 + B.rank)
    x = torch.empty((S, n), dtype=torch.int16, device=dev)
    t = torch.arange(n, device=dev, dtype=torch.float32) / 48000.0
    for s0 in range(0, S, 512):
        s1 = min(S, s0 + 512)
        f = (300.0 + 7.0 * (torch.arange(s_lo + s0, s_lo + s1, device=dev) % 64)).to(torch.float32)[:, None]
        v = 2000.0 * torch.randn((s1 - s0, n), generator=g, device=dev) + 2500.0 * torch.sin(2 * np.pi * f * t[None, :])
        x[s0:s1] = torch.clamp(torch.round(v), -32768, 32767).to(torch.int16)
        del v
    rng = np.random.default_rng(3 + B.rank)
    k = np.arange(512)
    h = rng.normal(0, 0.35, (S, 2, 512)) * np.exp(-k / 60.0)[None, None, :]
    h[:, :, :9] = 0.0
    h[:, :, 8] = 1.0
    h *= np.minimum(1.0, 3.0 / np.abs(h).sum(axis=2, keepdims=True))
    taps = np.concatenate([h, np.zeros((S, 2, 1))], axis=2)
    st = B.ctx.fastconv_state(p, S, taps)
    out = torch.empty((S, 2, n_out), dtype=torch.int16, device=dev)

    def run():
        st.reset()
        st.run(x, n, nb, out, n_out)
    ms, best = B.timed(run, warm=3, reps=5)
    alg = S * n * 2 + S * 2 * n_out * 2
    res = {"name": "fastconv", "workload": f"BASELINE.json configs[2]: {S_total} sources x {n / 48000:.0f} s @ 48 kHz x 512-tap HRIR pair -> binaural "
                                           f"(overlap-save, n_fft 1024, block 512; mode A: one output pair per source), {S} sources on this rank",
           "scaling": "strong", "unit": "Msamples/s (input samples)", "value": S_total * n / ms / 1e3, "ms": ms, "ms_best": best,
           "roofline": _roof("jdsp::fastconv_stream_kernel<512>", alg, ms, FLOP_PER_SAMPLE["fastconv"] * S * n,
                             "6 B per input sample (int16 in, two int16 ears out); filter spectra stay on chip")}
    if B.rank == 0 and not args.no_parity:
        run(); torch.cuda.synchronize()
        worst, flips, tot = 0, 0, 0
        for s in sorted({0, S // 3, S // 2, S - 1}):
            xs, got = x[s].cpu().numpy(), out[s].cpu().numpy()
            for ear in range(2):
                ref, _ = B.oracle.fastconv(xs, h[s, ear], Bk, 1, 1024)
                d = np.abs(got[ear].astype(int) - ref.astype(int))
                worst, flips, tot = max(worst, int(d.max())), flips + int((d > 0).sum()), tot + d.size
        res["parity"] = {"sources_checked": 4, "max_abs_lsb": worst, "flip_fraction": flips / tot}
    if B.world == 1 and not args.no_e2e:
        Se = min(S, 2048 if not args.quick else 256)
        ste = B.host_ctx().fastconv_state(p, Se, taps[:Se])
        h_in = torch.empty((Se, n), dtype=torch.int16).pin_memory()
        h_out = torch.empty((Se, 2, n_out), dtype=torch.int16).pin_memory()
        h_in.copy_(x[:Se].cpu())

        def call():
            ste.reset()
            assert ste.run_host(h_in, n, n, h_out, n_out) == n_out
        sec = B.wall(call)
        res["e2e"] = {"value": Se * n / sec / 1e6, "unit": "Msamples/s (input samples)", "h2d_bytes_per_step": Se * n * 2, "d2h_bytes_per_step": Se * 2 * n_out * 2,
                      "sources": Se, "api": "jdsp_fastconv_i16_host (pinned host in/out, chunks of sources over 3 CUDA streams)",
                      "equals_resident_path": bool(torch.equal(h_out[0], out[0].cpu()))}
        ste.close()
        del h_in, h_out
    st.close()
    del x, out
    B.free()
    if B.rank == 0 and B.world == 1 and not args.no_cpu:
        from jeicyboodsp_b200 import synth
        secs = 4.0
        nn = int(secs * 48000)

        def mk(d, c):
            fi, ft = os.path.join(d, f"in_{c}.wav"), os.path.join(d, f"taps_{c}.f64")
            with open(fi, "wb") as f:
                f.write(bytes(44) + synth.roundtrip_signal(nn, 48000.0, seed=c).tobytes())
            np.concatenate([synth.hrir_pair(c)[0], [0.0]]).tofile(ft)
            return fi, ft
        res["cpu_baseline"] = _cpu_program_baseline(
            "fastconv_bench", mk, lambda f: [f[0], f[0] + ".out", f[1]], nn,
            f"2 runs (one per ear) x {secs:.0f} s @ 48 kHz of Fast_Convolution_Based_3DAudio_Impl (block 512, n_fft 1024, 513 taps); input samples counted once", reps=2)
        if "value" in res["cpu_baseline"]:
            res["cpu_baseline"]["value"] /= 2            # both ears of a source = two program runs over the same input
            res["cpu_baseline"]["single_process_msamples_s"] /= 2
    return res


def cfg_mfcc(B: Bench):
    """BASELINE configs[3]: 25 ms / 10 ms frames, 512-pt FFT, 26 mel bands, 13 cepstra over 100 h of 16 kHz audio."""
    import numpy as np
    torch, dev, L, args = B.torch, B.dev, B.L, B.args
    from jeicyboodsp_b200 import synth
    from jeicyboodsp_b200.sharding import shard_range
    from oracle.oracle import MfccParams as OMP
    U_total, n = (36000, 160_000) if not args.quick else (2048, 160_000)
    u_lo, u_hi = shard_range(U_total, B.rank, B.world)
    U = u_hi - u_lo
    p = L.mfcc_params("bench")
    plan = B.ctx.mfcc_plan(p)
    nf = plan.n_frames(n)
    x = synth.denoise_streams_torch(U, n, dev, stream0=u_lo, sigma=25.0, seed=4)
    feat = torch.empty((U, nf, 13), dtype=torch.float32, device=dev)
    ms, best = B.timed(lambda: plan.run(x, n, U, n, feat, nf * 13), warm=3, reps=5)
    alg = U * n * 2 + U * nf * 13 * 4
    res = {"name": "mfcc", "workload": f"BASELINE.json configs[3]: {U_total} utterances x 10 s = {U_total * n / FS / 3600:.0f} h @ 16 kHz, frames 400 / hop 160, "
                                       f"512-pt FFT, 26 mel bands, 13 cepstra (f32 rows), {U} utterances on this rank",
           "scaling": "strong", "unit": "Msamples/s", "value": U_total * n / ms / 1e3, "frames_per_s": U_total * nf / ms * 1e3, "ms": ms, "ms_best": best,
           "roofline": _roof("jdsp::mfcc_kernel<256,13>", alg, ms, FLOP_PER_SAMPLE["mfcc"] * U * n, "2.325 B per sample (int16 in + 13 f32 per 160 samples)")}
    if B.rank == 0 and not args.no_parity:
        torch.cuda.synchronize()
        worst = 0.0
        for u in sorted({0, U // 3, U // 2, U - 1}):
            ref = B.oracle.mfcc_frames(x[u].cpu().numpy(), OMP.preset("bench"))
            worst = max(worst, float(np.abs(feat[u].cpu().numpy() - ref).max() / np.abs(ref).max()))
        res["parity"] = {"utterances_checked": 4, "max_err_rel_peak": worst, "tolerance": 1e-4}
    if B.world > 1:
        # one feature matrix on every rank: NCCL all-gather of the per-GPU blocks, after the kernel and overlapped with it
        dist = B.dist
        Umax = max(shard_range(U_total, r, B.world)[1] - shard_range(U_total, r, B.world)[0] for r in range(B.world))
        pad = torch.zeros((Umax, nf, 13), dtype=torch.float32, device=dev)
        full = torch.empty((B.world * Umax, nf, 13), dtype=torch.float32, device=dev)

        def serial():
            plan.run(x, n, U, n, pad, nf * 13)
            dist.all_gather_into_tensor(full, pad)
        ms_serial, _ = B.timed(serial, warm=2, reps=4)
        NCH = 8
        cu = (Umax + NCH - 1) // NCH
        fullc = torch.empty((NCH, B.world, cu, nf, 13), dtype=torch.float32, device=dev)
        padc = torch.zeros((NCH, cu, nf, 13), dtype=torch.float32, device=dev)

        def overlapped():
            hs = []
            for ch in range(NCH):
                a, b = ch * cu, min(U, (ch + 1) * cu)
                if b > a:
                    plan.run(x[a:b], n, b - a, n, padc[ch], nf * 13)
                hs.append(dist.all_gather_into_tensor(fullc[ch].view(B.world * cu, nf, 13), padc[ch], async_op=True))
            for hnd in hs:
                hnd.wait()
        ms_over, _ = B.timed(overlapped, warm=2, reps=4)
        gathered_ok = bool(torch.equal(fullc[0, B.rank, : min(cu, U)], feat[: min(cu, U)]))
        # the same result with NO collective: the scatter form of the kernel writes every feature row into every rank's copy of the matrix
        # (peer memory over NVLink, jeicyboodsp_b200.sharding.PeerMatrix); a one-element all-reduce inside the timed region stands for "all
        # ranks' rows have landed", which is what the all-gather's completion means
        from jeicyboodsp_b200.sharding import PeerMatrix
        serial()                                       # `full` = the NCCL result to compare with
        tiny = torch.zeros(1, dtype=torch.float32, device=dev)
        pm, pm_err = None, ""
        try:
            pm = PeerMatrix(B.ctx, U_total, nf * 13)
        except Exception as e:   # noqa: BLE001 - a box without peer access between its GPUs: the NCCL lines above stand alone
            pm_err = f"{type(e).__name__}: {e}"[:200]
        ms_fused, fused_ok = None, False
        all_have = B.min_ranks_bool(pm is not None)
        if all_have:
            dests = pm.dests(u_lo)

            def fused():
                plan.run_scatter(x, n, U, n, dests, nf * 13)
                dist.all_reduce(tiny)
            ms_fused, _ = B.timed(fused, warm=2, reps=4)
            torch.cuda.synchronize()
            B.barrier()
            M = pm.tensor().view(U_total, nf, 13)
            fused_ok = True
            for r in range(B.world):
                lo, hi = shard_range(U_total, r, B.world)
                fused_ok = fused_ok and bool(torch.equal(M[lo:hi], full[r * Umax: r * Umax + (hi - lo)]))
            del M
        if pm is not None:
            pm.close(collective=all_have)
        # ... and through ONE NVLink multicast address: each row leaves the GPU once, the NVSwitch replicates it into every copy
        mc = {"available": False}
        try:
            from jeicyboodsp_b200.sharding import MulticastMatrix
            mm = MulticastMatrix(U_total, nf * 13, dev)
            if mm.available():
                mcd = mm.dest(u_lo)

                def fused_mc():
                    plan.run_multicast(x, n, U, n, mcd, nf * 13)
                    dist.all_reduce(tiny)
                ms_mc, _ = B.timed(fused_mc, warm=2, reps=4)
                torch.cuda.synchronize()
                B.barrier()
                Mm = mm.tensor().view(U_total, nf, 13)
                mc_ok = True
                for r in range(B.world):
                    lo, hi = shard_range(U_total, r, B.world)
                    mc_ok = mc_ok and bool(torch.equal(Mm[lo:hi], full[r * Umax: r * Umax + (hi - lo)]))
                mc = {"available": True, "ms": ms_mc, "equals_nccl_result_on_every_rank": bool(B.min_ranks_bool(mc_ok)),
                      "msamples_s": U_total * n / ms_mc / 1e3,
                      "how": "jdsp_mfcc_frames_i16_multicast_dev: multimem.st to the multicast address of torch symmetric memory (NVSwitch replicates "
                             "each 256-byte warp store into all ranks' matrices), then a 1-element all-reduce"}
                del Mm
            del mm
        except Exception as e:   # noqa: BLE001 - no symmetric memory / multicast on this platform: the unicast scatter above stands
            mc = {"available": False, "why": f"{type(e).__name__}: {e}"[:200]}
        res["gather"] = {"collective": "NCCL all_gather_into_tensor of [utterances/N, 998, 13] f32 blocks -> one matrix on every rank",
                         "ms_kernel_only": ms, "ms_kernel_then_gather": ms_serial, "ms_chunked_overlap": ms_over, "chunks": NCH,
                         "bytes_received_per_rank": int((B.world - 1) * Umax * nf * 13 * 4), "own_block_intact": gathered_ok,
                         "msamples_s_with_gather": U_total * n / ms_over / 1e3,
                         "ms_fused_scatter": ms_fused, "fused_scatter_equals_nccl_result_on_every_rank": bool(B.min_ranks_bool(fused_ok)),
                         "msamples_s_fused_scatter": (U_total * n / ms_fused / 1e3) if ms_fused else None, "fused_scatter_error": pm_err or None,
                         "fused_multicast": mc,
                         "fused_scatter": "jdsp_mfcc_frames_i16_scatter_dev: the kernel's feature rows written to all ranks' matrices through "
                                          "CUDA-IPC peer mappings over NVLink (256-byte runs per warp store), then a 1-element all-reduce; no all-gather"}
        del pad, full, fullc, padc
    if B.world == 1 and not args.no_e2e:
        Ue = min(U, 4500 if not args.quick else 512)
        h_in = torch.empty((Ue, n), dtype=torch.int16).pin_memory()
        h_out = torch.empty((Ue, nf, 13), dtype=torch.float32).pin_memory()
        h_in.copy_(x[:Ue].cpu())
        eplan = B.host_ctx().mfcc_plan(p)
        sec = B.wall(lambda: eplan.run_host(h_in, n, Ue, n, h_out, nf * 13))
        res["e2e"] = {"value": Ue * n / sec / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": Ue * n * 2, "d2h_bytes_per_step": Ue * nf * 13 * 4,
                      "utterances": Ue, "api": "jdsp_mfcc_frames_i16 (pinned host in/out, chunks of utterances over 3 CUDA streams)",
                      "equals_resident_path": bool(torch.equal(h_out[0], feat[0].cpu()))}
        eplan.close()
        del h_in, h_out
    plan.close()
    del x, feat
    B.free()
    if B.rank == 0 and B.world == 1 and not args.no_cpu:
        secs = 60.0
        nn = int(secs * FS)

        def mk(d, c):
            fi, fl = os.path.join(d, f"in_{c}.wav"), os.path.join(d, f"list_{c}.txt")
            with open(fi, "wb") as f:
                f.write(bytes(44) + synth.mfcc_utterance(c, nn).tobytes())
            with open(fl, "w") as f:
                f.write(f"{fi} {fi}.mfc")          # no trailing newline
            return (fl,)
        res["cpu_baseline"] = _cpu_program_baseline(
            "mfcc_mid", mk, lambda f: [f[0]], nn,
            f"{secs:.0f} s @ 16 kHz of MFCCFeatureExtraction_auto_version1 at the `mid` preset (512-sample frames, hop 256, 26 mel, 13 cepstra: the "
            "closest framing the program's defines can express; the bench framing 400 / 160 makes 1.6x as many frames per sample)")
    return res


def _committed_traffic(S, n):
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from capture_traffic import source_hash
        doc = json.load(open(os.path.join(ROOT, "profiles", "round2", "traffic_denoise.json")))
        if doc["kernel_sources_sha256"] != source_hash():
            return None, "profiles/round2/traffic_denoise.json is stale (kernel sources changed since the capture)"
        if doc["streams"] != S or doc["samples_per_stream"] != n:
            return None, "profiles/round2/traffic_denoise.json was captured at another launch size"
        return doc["dram_bytes_per_launch"], ("ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of one SS and one Wiener launch at this size "
                                              "(profiles/round2/traffic_denoise.json, kernel sources sha256 " + doc["kernel_sources_sha256"][:12] + ")")
    except Exception as e:   # noqa: BLE001 - a missing capture is not an error of the run
        return None, f"no committed capture ({type(e).__name__})"


def cfg_sweep(B: Bench):
    """BASELINE configs[4]: batched complex64 FFT, N = 2^8 .. 2^16, 2^29 points = 4 GiB in + 4 GiB out per size."""
    import numpy as np
    torch, dev, args = B.torch, B.dev, B.args
    total_all = (1 << 29) if not args.quick else (1 << 25)
    total = total_all // B.world                         # batch rows are independent: contiguous shards
    peak, _ = _peaks()
    xy = torch.empty(2 * total, dtype=torch.complex64, device=dev)   # input then output, one buffer size apart
    x, y = xy[:total], xy[total:]
    g = torch.Generator(device=dev); g.manual_seed(5 + B.rank)
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    rows, fr = [], []
    for lg in range(8, 17):
        n = 1 << lg
        batch = total // n
        row = {"n": n, "batch_per_rank": batch}
        for fwd, nm in ((True, "fwd"), (False, "inv")):
            # every (size, direction) is its own "8 GB run": 0.1 s of idle device, 3 warm-up launches, 6 timed ones -- the clock state the HBM
            # peak itself was measured in.  Timed back to back instead, the sizes whose CTAs run load / transform / store phases one after
            # the other (N >= 4096) lose 3-12 % within a second of sustained load (profiles/round2/fft_offset_vs_time.txt); both are reported.
            ms, best = B.timed(lambda: B.ctx.fft_c2c_f32(x, y, n, batch, fwd), warm=3, reps=6, idle_s=0.1)
            row[f"{nm}_ms"] = ms
            row[f"{nm}_frac_hbm"] = total * 16 / ms / 1e6 / peak
            row[f"{nm}_frac_hbm_best"] = total * 16 / best / 1e6 / peak
            row[f"{nm}_fp32_frac"] = 5.0 * lg * total / (ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS
            fr.append(row[f"{nm}_frac_hbm"])
        row["gpoints_s"] = total_all / row["fwd_ms"] / 1e6
        if B.rank == 0 and not args.no_parity:
            B.ctx.fft_c2c_f32(x, y, n, batch, True)
            torch.cuda.synchronize()
            worst = 0.0
            for r in np.random.default_rng(n).integers(0, batch, 4):
                zi = x[r * n:(r + 1) * n].cpu().numpy().astype(np.complex128)
                ref = B.oracle.fftprocess(zi, True) if n <= 32768 else np.fft.fft(zi)   # the reference breaks at 2^16 (short indices)
                worst = max(worst, float(np.abs(y[r * n:(r + 1) * n].cpu().numpy() - ref).max() / np.abs(ref).max()))
            row["parity_max_rel"] = worst
        rows.append(row)
    for row in rows:   # the same launches back to back, all sizes in one go (sustained load)
        n, batch = row["n"], row["batch_per_rank"]
        for fwd, nm in ((True, "fwd"), (False, "inv")):
            ms, _ = B.timed(lambda: B.ctx.fft_c2c_f32(x, y, n, batch, fwd), warm=3, reps=6)
            row[f"{nm}_frac_hbm_back_to_back"] = total * 16 / ms / 1e6 / peak
    worst_row = min(rows, key=lambda r: min(r["fwd_frac_hbm"], r["inv_frac_hbm"]))
    res = {"name": "fft_sweep", "workload": f"BASELINE.json configs[4]: batched complex64 FFT, N = 2^8..2^16, {total_all} points (4 GiB in + 4 GiB out) per size, "
                                            f"forward and inverse, {total} points on this rank", "scaling": "strong", "unit": "Gpoints/s",
           "value": statistics.mean(r["gpoints_s"] for r in rows), "sizes": rows,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "frac_mean": statistics.mean(fr), "frac_min": min(fr), "frac_min_n": worst_row["n"],
                        "achieved_mean": statistics.mean(fr) * peak, "algorithmic_bytes_per_launch": total * 16,
                        "frac_mean_back_to_back": statistics.mean(r[f"{d}_frac_hbm_back_to_back"] for r in rows for d in ("fwd", "inv")),
                        "note": "16 B per point (complex64 in + out); the one HBM-bound config; every size and direction timed as its own run "
                                "(0.1 s idle, 3 warm-up launches, mean of 6), *_back_to_back = all sizes in one uninterrupted go"}}
    if B.world == 1 and not args.no_e2e:
        tot_e = min(total, 1 << 26)
        h_in = torch.empty(tot_e, dtype=torch.complex64).pin_memory()
        h_out = torch.empty(tot_e, dtype=torch.complex64).pin_memory()
        h_in.copy_(x[:tot_e].cpu())
        n_e = 4096
        sec = B.wall(lambda: B.host_ctx().fft_c2c_f32_host(h_in, h_out, n_e, tot_e // n_e, True))
        B.ctx.fft_c2c_f32(x, y, n_e, tot_e // n_e, True)
        torch.cuda.synchronize()
        res["e2e"] = {"value": tot_e / sec / 1e9, "unit": "Gpoints/s", "n": n_e, "h2d_bytes_per_step": tot_e * 8, "d2h_bytes_per_step": tot_e * 8,
                      "api": "jdsp_fft_c2c_f32_host (pinned host in/out, chunks of transforms over 3 CUDA streams)",
                      "equals_resident_path": bool(torch.equal(h_out[:n_e], y[:n_e].cpu()))}
        del h_in, h_out
    del x, y, xy
    B.free()
    if B.rank == 0 and B.world == 1 and not args.no_cpu:
        from oracle.oracle import RefPrograms
        ref = RefPrograms()
        if os.path.exists(os.path.join(ref.dir, "libfftprocess_4096.so")):
            z = (np.random.default_rng(5).uniform(-1, 1, (48, 4096)) + 0j)
            ref.fftprocess(z[:2], True)
            t0 = time.perf_counter()
            ref.fftprocess(z, True)
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": z.size / dt / 1e9, "unit": "Gpoints/s", "cores": 1, "kind": "reference",
                                   "sample": "48 forward transforms of 4096 points through the reference's own FFTProcess (FFTAlgorithm_ver2.cpp built with BLOCK_LEN 4096), one thread; "
                                             "the reference has no batched or multi-threaded FFT driver"}
        else:
            res["cpu_baseline"] = {"unavailable": "oracle/_ref/libfftprocess_4096.so not built"}
    return res


def cfg_roundtrip(B: Bench):
    """BASELINE configs[0]: FFTAlgorithm_ver2's FFT -> IFFT round trip at 1024 points over the 10 s 16 kHz signal; the single
    stream (latency) and 65 536 replicas of it (roofline)."""
    import numpy as np
    torch, dev, args = B.torch, B.dev, B.args
    from jeicyboodsp_b200 import synth
    n_fft = 1024
    sig = synth.roundtrip_signal(160_000)
    nb = -(-len(sig) // n_fft)
    row = nb * n_fft
    pad = np.zeros(row, np.int16); pad[: len(sig)] = sig
    pad[len(sig):] = pad[len(sig) - n_fft: row - n_fft]          # the program's stale tail (FFTAlgorithm_ver2.cpp:64)
    R_total = 65536 if not args.quick else 2048
    R = R_total // B.world
    one = torch.from_numpy(pad).to(dev)
    x = one.unsqueeze(0).repeat(R, 1).contiguous()
    y = torch.empty_like(x)
    ms1, best1 = B.timed(lambda: B.ctx.roundtrip_dev(one, row, y, row, None, 0, n_fft, 1, nb), warm=3, reps=10)
    ms, best = B.timed(lambda: B.ctx.roundtrip_dev(x, row, y, row, None, 0, n_fft, R, nb), warm=3, reps=5)
    alg = R * row * 4
    res = {"name": "roundtrip", "workload": f"BASELINE.json configs[0]: FFTAlgorithm_ver2 1024-pt FFT -> IFFT round trip over the 10 s 16 kHz signal, "
                                            f"{R_total} replicas ({R} on this rank) for the roofline; the single stream for latency",
           "scaling": "strong", "unit": "Msamples/s", "value": R_total * row / ms / 1e3, "ms": ms, "ms_best": best,
           "single_stream_us": ms1 * 1e3, "single_stream_msamples_s": row / ms1 / 1e3,
           "roofline": _roof("jdsp::roundtrip_warp_kernel<1024>", alg, ms, FLOP_PER_SAMPLE["roundtrip"] * R * row, "4 B per sample (int16 in + int16 out)")}
    if B.rank == 0 and not args.no_parity:
        f32 = torch.empty((1, row), dtype=torch.float32, device=dev)
        B.ctx.roundtrip_dev(one, row, y, row, f32, row, n_fft, 1, nb)
        torch.cuda.synchronize()
        ref_i16, ref_f64 = B.oracle.roundtrip(pad, n_fft)
        res["parity"] = {"precast_max_err_rel_peak": float(np.abs(f32[0].cpu().numpy() - ref_f64).max() / np.abs(ref_f64).max()),
                         "i16_max_abs_lsb": int(np.abs(y[0].cpu().numpy().astype(int) - ref_i16.astype(int)).max()), "tolerance": 1e-4,
                         "note": "bit-exact int16 is impossible by construction: the reference's own (short) cast truncates k +- 1e-11 (SURVEY 0.3-1)"}
    if B.world == 1 and not args.no_e2e:
        Re = min(R, 8192 if not args.quick else 512)
        h_in = torch.empty((Re, row), dtype=torch.int16).pin_memory()
        h_out = torch.empty((Re, row), dtype=torch.int16).pin_memory()
        h_in.copy_(x[:Re].cpu())
        sec = B.wall(lambda: B.host_ctx().roundtrip_batch_raw(h_in, row, Re, row, n_fft, h_out, row))
        B.ctx.roundtrip_dev(x, row, y, row, None, 0, n_fft, Re, nb)
        torch.cuda.synchronize()
        res["e2e"] = {"value": Re * row / sec / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": Re * row * 2, "d2h_bytes_per_step": Re * row * 2, "replicas": Re,
                      "api": "jdsp_roundtrip_batch_i16 (pinned host in/out, chunks of streams over 3 CUDA streams)",
                      "equals_resident_path": bool(torch.equal(h_out[0], y[0].cpu()))}
        del h_in, h_out
    del x, y
    B.free()
    if B.rank == 0 and B.world == 1 and not args.no_cpu:
        def mk(d, c):
            fi = os.path.join(d, f"in_{c}.wav")
            with open(fi, "wb") as f:
                f.write(bytes(44) + np.roll(sig, 131 * c).tobytes())
            return (fi,)
        res["cpu_baseline"] = _cpu_program_baseline("fft_roundtrip_1024", mk, lambda f: [f[0], f[0] + ".out"], len(sig),
                                                    "20 runs over the 10 s 16 kHz signal of FFTAlgorithm_ver2 built with BLOCK_LEN 1024 (its own FFTProcess, no FFTW)", reps=20,
                                                    fftw_shim=False)
    return res


# ---------------------------------------------------------------------------------------------------------
def main_gpu(args):
    import numpy as np
    import torch

    from jeicyboodsp_b200 import synth
    from jeicyboodsp_b200.binding import SS, WIENER

    # libraries (NCCL prints its version banner) must not pollute stdout: the contract is ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    B = Bench(args)
    world, rank, local, dev, L, ctx = B.world, B.rank, B.local, B.dev, B.L, B.ctx
    only = set(args.only.split(",")) if args.only else {"denoise", "fastconv", "mfcc", "sweep", "roundtrip"}

    S, n = args.streams, int(args.seconds * FS)
    nb = n // HOP
    n = nb * HOP
    n_out = (nb - 2) * HOP
    x = synth.denoise_streams_torch(S, n, dev, stream0=rank * S)
    out = torch.empty((S, n_out), dtype=torch.int16, device=dev)
    params = {m: L.denoise_params("bench", m) for m in (SS, WIENER)}
    states = {m: ctx.denoise_state(params[m], S) for m in (SS, WIENER)}
    alg_bytes = S * (n + n_out) * 2  # one int16 read and one int16 write per sample, per launch

    def step(events=None):
        for m in (SS, WIENER):
            states[m].reset()
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            states[m].run(x, n, nb, out, n_out)
            if events is not None:
                e1.record()
                events.append((m, e0, e1))

    for _ in range(args.warmup):
        step()
    B.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches()
    evs = []
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step(evs)
    t_end.record()
    B.barrier()
    launches = ctx.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = B.max_ranks(t_start.elapsed_time(t_end))
    ms_per_step = ms_total / args.steps
    samples_per_step = world * S * n * 2
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6
    kern_ms = [e0.elapsed_time(e1) for (_, e0, e1) in evs]
    per_mode = {("ss" if m == SS else "wiener"): statistics.mean([e0.elapsed_time(e1) for (mm, e0, e1) in evs if mm == m])
                for m in (SS, WIENER)}
    avg_kernel_ms = statistics.mean(kern_ms)
    roofline = _roof("jdsp::denoise_stream_kernel<256,MODE,16>", alg_bytes, avg_kernel_ms, FLOP_PER_SAMPLE["denoise"] * S * n,
                     "4 B/sample (int16 in + int16 out); one half warp per stream; the fp32 transforms and their shared-memory traffic, not HBM, bound it (DESIGN.md)")
    roofline["per_mode_ms"] = per_mode
    # dram__bytes of one launch cannot be measured inside an unprofiled run: it comes from a committed ncu capture of the same launch size
    # (tools/capture_traffic.py) and is reported only while the kernel sources still hash to what was captured; otherwise null
    roofline["traffic"], roofline["traffic_source"] = _committed_traffic(S, n)

    # ---- parity spot check on the very data that was timed (8 streams through the oracle) -------------------
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle.oracle import DenoiseParams as ODP
        pick = sorted(set([0, 1, S // 2 - 1, S - 1] + [int(v) for v in np.random.default_rng(2).integers(0, S, 4)]))
        worst, flips, total, pubs, ambiguous = 0, 0, 0, [], 0
        for m in (SS, WIENER):
            states[m].reset()
            states[m].run(x, n, nb, out, n_out)
            got = out[pick].cpu().numpy()
            for i, s in enumerate(pick):
                r = B.oracle.denoise(x[s].cpu().numpy(), ODP.preset("bench", m))
                # Blocks at zcr == thr-1 with low energy are undecidable in the REFERENCE BINARY (it reads one element past its VAD
                # buffer); the oracle and the kernels both define that element as 0, so they are compared on
                # every block.  The count is reported because parity against the compiled program (tests/) has to skip such streams.
                ambiguous += int(np.sum((r.zcr == params[m].zcr_thr - 1) & (r.energy <= params[m].energy_thr)))
                d = np.abs(got[i].astype(np.int64) - r.out.astype(np.int64))
                worst, flips, total = max(worst, int(d.max())), flips + int((d > 0).sum()), total + d.size
                pubs.append(len(r.publish))
        parity = {"streams_checked": len(pick), "max_abs_lsb": worst, "flip_fraction": flips / max(total, 1), "oracle_publishes_min": min(pubs),
                  "blocks_ambiguous_in_the_reference_binary": ambiguous}

    # ---- end to end through the host-buffer C-ABI call, and the raw copies alone as a control ----------------------
    e2e = None
    if not args.no_e2e:
        with open("/proc/meminfo") as f:
            avail_kb = int([l for l in f if l.startswith("MemAvailable")][0].split()[1])
        need = 2 * S * n * 2 * world * 1.3
        Se = S if avail_kb * 1024 > need else max(64, S // 8)
        h_in = torch.empty((Se, n), dtype=torch.int16).pin_memory()
        h_out = torch.empty((Se, n_out), dtype=torch.int16).pin_memory()
        h_in.copy_(x[:Se].cpu())
        ectx = B.host_ctx()
        k_e2e = max(1, min(args.steps, 3))

        def e2e_step():
            for m in (SS, WIENER):
                got = ectx.denoise_host_raw(params[m], h_in, n, Se, n, h_out, n_out)
                assert got == n_out
        sec = B.wall(e2e_step, reps=k_e2e)
        e2e_val = world * Se * n * 2 / sec / 1e6
        if rank == 0 and parity is not None:
            states[WIENER].reset()
            states[WIENER].run(x, n, nb, out, n_out)
            torch.cuda.synchronize()
            parity["e2e_equals_resident_path"] = bool(torch.equal(h_out[0], out[0].cpu()) and torch.equal(h_out[Se - 1], out[Se - 1].cpu()))
        # control: the same bytes, pinned H2D and D2H at once on two streams, no kernels, all ranks at the same time
        d_a, d_b = torch.empty((Se, n), dtype=torch.int16, device=dev), out[:Se]
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def copies():
            for _ in range(2):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
            s1.synchronize(); s2.synchronize()
        sec_c = B.wall(copies, reps=2)
        bytes_step = 2 * Se * (n + n_out) * 2
        e2e = {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": 2 * Se * n * 2, "d2h_bytes_per_step": 2 * Se * n_out * 2,
               "streams_per_gpu": Se, "steps": k_e2e, "timer": "host wall clock around the blocking C-ABI calls, max over ranks",
               "gbs_per_rank": bytes_step / sec / 1e9, "copy_only_gbs_per_rank": bytes_step / sec_c / 1e9,
               "copy_only_msamples_s": world * Se * n * 2 / sec_c / 1e6,
               "copy_only_note": "the same pinned buffers copied H2D and D2H concurrently on two streams with no kernels, all ranks at once, same timer: "
                                 "the ceiling the platform's host memory / PCIe path gives this API",
               "api": "jdsp_denoise_i16 (pinned host in/out; chunks of TIME, all streams per chunk, H2D / kernel / D2H over 3 CUDA streams)"}
        del h_in, h_out, d_a

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        agg, cores, kind, what, single, _ = run_reference_cpu(args.ref_streams_per_core, args.ref_seconds)
        cpu = {"value": agg, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": what, "single_process_msamples_s": single}
    for st in states.values():
        st.close()
    del x, out
    B.free()

    # ---- the other BASELINE configs -------------------------------------------------------------------------------
    configs = []
    launches_cfg0 = ctx.kernel_launches()
    for name, fn in (("fastconv", cfg_fastconv), ("mfcc", cfg_mfcc), ("sweep", cfg_sweep), ("roundtrip", cfg_roundtrip)):
        if name in only:
            t0 = time.perf_counter()
            try:
                r = fn(B)
            except Exception as ex:  # a failing extra config must not take the headline line with it; it is reported, not hidden
                r = {"name": name, "error": f"{type(ex).__name__}: {ex}"}
                B.free()
            r["wall_s"] = time.perf_counter() - t0
            configs.append(r)
    launches_cfg = ctx.kernel_launches() - launches_cfg0

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "denoise: SS + Wiener, 512-pt Hann STFT, hop 256, 16 kHz speech+AWGN, "
                                   f"{S} streams x {n / FS:.0f} s per GPU (BASELINE.json configs[1])",
                       "streams_per_gpu": S, "samples_per_stream": n, "n_fft": N_FFT, "hop": HOP, "zcr_thr": 64,
                       "passes_per_step": ["spectral_subtraction", "wiener"],
                       "l2": f"inputs {S * n * 2 / 1e9:.2f} GB per pass >> 126 MB L2 (no flush needed)",
                       "frames_per_s": value * 1e6 / HOP},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity,
            "configs": configs, "gpu_launches_configs": launches_cfg,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if B.ectx is not None:
        B.ectx.close()
    if world > 1:
        B.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU")
    ap.add_argument("--seconds", type=float, default=60.0, help="seconds of 16 kHz audio per stream")
    ap.add_argument("--only", default="", help="comma list of extra configs to run: fastconv,mfcc,sweep,roundtrip (default all)")
    ap.add_argument("--quick", action="store_true", help="small extra configs (smoke run of the harness, not a measurement)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--ref-streams-per-core", type=int, default=2)
    ap.add_argument("--ref-seconds", type=float, default=60.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing hygiene: at least three warm-up steps
    return main_reference(args) if args.impl == "reference" else main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
