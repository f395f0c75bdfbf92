"""Oracle restatement vs the committed fixtures produced by the reference's own code
(tests/golden/make_golden.py ran the unmodified programs of oracle/_ref)."""
import os

import numpy as np
import pytest

from oracle.oracle import DenoiseParams, MfccParams

G = os.path.join(os.path.dirname(__file__), "golden")


def test_fft_golden(oracle):
    g = np.load(os.path.join(G, "fft.npz"))
    for n in (256, 512, 1024, 4096, 32768):
        assert np.array_equal(oracle.fftprocess(g[f"in_{n}"], True), g[f"fwd_{n}"])
        assert np.array_equal(oracle.fftprocess(g[f"in_{n}"], False), g[f"inv_{n}"])
    assert np.array_equal(oracle.bitrev_table(512), g["bitrev512"].astype(np.int32))
    assert np.array_equal(oracle.bitrev_table(32768), g["bitrev32768"].astype(np.int32))
    assert np.array_equal(oracle.dftprocess(g["pcm"][:512]), g["dft512"])
    for n in (512, 1024):
        assert np.array_equal(oracle.roundtrip(g["pcm"], n)[0], g[f"rt{n}"])


@pytest.mark.parametrize("preset", ["ref", "bench"])
def test_denoise_golden(oracle, preset):
    g = np.load(os.path.join(G, "denoise.npz"))
    for stream in (3, 17):
        for mode, nm in ((0, "ss"), (1, "wiener")):
            res = oracle.denoise(g[f"pcm_{stream}"], DenoiseParams.preset(preset, mode))
            assert len(res.publish) > 0
            assert np.array_equal(res.out, g[f"{nm}_{preset}_{stream}"])
        assert set(np.unique(g[f"zcr_{preset}_{stream}"] - res.zcr)) <= {0, 1}


def test_fastconv_golden(oracle):
    g = np.load(os.path.join(G, "fastconv.npz"))
    for ear in range(2):
        got, _ = oracle.fastconv(g["pcm_bench"], g["hrir_bench"][ear], 512, 1, 1024)
        assert np.array_equal(got, g[f"out_bench_ear{ear}"])
    taps = np.zeros(7169)
    taps[g["ref_taps_idx"]] = g["ref_taps_val"]
    got, _ = oracle.fastconv(g["pcm_ref"], taps, 1024, 7, 8192)
    assert np.array_equal(got, g["out_ref"])


def test_mfcc_golden(oracle):
    g = np.load(os.path.join(G, "mfcc.npz"))
    for preset in ("ref", "mid"):
        got = oracle.mfcc_program(g["pcm"], MfccParams.preset(preset))
        assert got.shape == g[preset].shape
        assert np.abs(got - g[preset]).max() < 1e-9


def test_pitch_golden(oracle):
    g = np.load(os.path.join(G, "pitch.npz"))
    for stream in (3, 17):
        for exact in (False, True):
            arg, mx = oracle.pitch(g[f"pcm_{stream}"], exact=exact)
            assert np.array_equal(arg, g[f"arg_{stream}"])
            assert np.allclose(mx, g[f"rmax_{stream}"], rtol=0, atol=1e-6 + 1e-9 * np.abs(mx).max())


def test_mvdr_golden(oracle):
    """BeamForming_MVDR_ver1 outputs of the unmodified program (Eigen served by oracle/eigen_shim): bit-exact."""
    g = np.load(os.path.join(G, "mvdr.npz"))
    for stream in (3, 17):
        out, pre, corr, vad = oracle.mvdr(g[f"left_{stream}"], g[f"right_{stream}"])
        assert (vad == 0).sum() > 2 and corr[-1, 0] > 0 and corr[-1, 3] > 0, "spatial matrix never estimated: vacuous"
        assert np.array_equal(out, g[f"out_{stream}"])
