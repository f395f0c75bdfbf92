"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference programs (oracle/_ref, built by
oracle/build.sh from /root/reference) on small synthetic inputs.  Run in the build container:

    python tests/golden/make_golden.py

The fixtures pin both the oracle restatement (CPU tests) and the CUDA path (GPU tests) to outputs of the
reference's own code; /root/reference does not exist on the GPU box, so they are committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from jeicyboodsp_b200 import synth  # noqa: E402
from oracle.oracle import RefPrograms  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main() -> None:
    r = RefPrograms()
    assert r.available(), "oracle/_ref missing: run oracle/build.sh where /root/reference exists"
    # --- FFTProcess / round trip -----------------------------------------------------------------
    x = synth.roundtrip_signal(12_000 + 100)          # not a multiple of the block: exercises the stale tail
    rng = np.random.default_rng(11)
    fft = {}
    for n in (256, 512, 1024, 4096, 32768):
        z = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n))
        fft[f"in_{n}"] = z
        fft[f"fwd_{n}"] = r.fftprocess(z, True)
        fft[f"inv_{n}"] = r.fftprocess(z, False)
    np.savez_compressed(os.path.join(OUT, "fft.npz"), pcm=x, rt512=r.roundtrip(x, 512), rt1024=r.roundtrip(x, 1024),
                        dft512=r.dftprocess(x[:512]), bitrev512=r.bitrev_table(512), bitrev32768=r.bitrev_table(32768),
                        **fft)
    # --- denoise ------------------------------------------------------------------------------------
    d = {}
    for stream in (3, 17):
        xs = synth.denoise_stream(stream, 48_000 + 333)
        d[f"pcm_{stream}"] = xs
        for preset in ("ref", "bench"):
            for mode, nm in ((0, "ss"), (1, "wiener")):
                out, en, zc = r.denoise(xs, preset, mode, want_vad=True)
                d[f"{nm}_{preset}_{stream}"] = out
                if mode == 0:
                    d[f"energy_{preset}_{stream}"] = en
                    d[f"zcr_{preset}_{stream}"] = zc
    np.savez_compressed(os.path.join(OUT, "denoise.npz"), **d)
    # --- fast convolution -----------------------------------------------------------------------------
    f = {}
    xs = synth.fastconv_source(5, 24_000 + 77)
    h = synth.hrir_pair(5)
    f["pcm_bench"], f["hrir_bench"] = xs, h
    for ear in range(2):
        f[f"out_bench_ear{ear}"] = r.fastconv(xs, "bench", np.concatenate([h[ear], [0.0]]))
    xr = synth.fastconv_source(2, 1024 * 14 + 500)
    f["pcm_ref"] = xr
    f["out_ref"] = r.fastconv(xr, "ref")               # the program's own 7169-tap room response
    txt = open(os.path.join(os.environ.get("JDSP_REFERENCE_DIR", "/root/reference"), "FilterCoefficient.h")).read()
    body = txt[txt.index("{") + 1: txt.rindex("}")]
    taps = np.array([float(v) for v in body.replace("\n", " ").split(",") if v.strip()])
    nz = np.nonzero(taps)[0]
    f["ref_taps_idx"], f["ref_taps_val"] = nz.astype(np.int32), taps[nz]   # 69 non-zero taps of FilterCoefficient.h:4
    np.savez_compressed(os.path.join(OUT, "fastconv.npz"), **f)
    # --- MFCC --------------------------------------------------------------------------------------------
    m = {}
    xu = synth.mfcc_utterance(1, 24_000 + 55)
    m["pcm"] = xu
    m["ref"] = r.mfcc(xu, "ref", 12)
    m["mid"] = r.mfcc(xu, "mid", 13)
    np.savez_compressed(os.path.join(OUT, "mfcc.npz"), **m)
    # --- pitch (PitchEstimation_method1) --------------------------------------------
    pt = {}
    for stream in (3, 17):
        xs = synth.denoise_stream(stream, 40_000 + 211)   # gated harmonic "speech": voiced and noise-only blocks, stale tail
        pt[f"pcm_{stream}"] = xs
        pt[f"arg_{stream}"], pt[f"rmax_{stream}"] = r.pitch(xs)
    np.savez_compressed(os.path.join(OUT, "pitch.npz"), **pt)
    # --- MVDR (BeamForming_MVDR_ver1; Eigen served by oracle/eigen_shim) -------------
    mv = {}
    for stream in (3, 17):
        xl, xr = synth.mvdr_pair(stream, 36_000 + 301)
        mv[f"left_{stream}"], mv[f"right_{stream}"] = xl, xr
        mv[f"out_{stream}"] = r.mvdr(xl, xr)
    np.savez_compressed(os.path.join(OUT, "mvdr.npz"), **mv)
    for fn in sorted(os.listdir(OUT)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(OUT, fn)), "bytes")


if __name__ == "__main__":
    main()
