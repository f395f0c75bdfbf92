"""Pins the oracle restatement (oracle/jdsp_oracle.c) against the UNMODIFIED reference programs compiled
by oracle/build.sh (oracle/_ref).  Skipped where those binaries are absent; the committed fixtures in
tests/golden/ (made from the same binaries) cover that case in test_oracle_golden.py."""
import numpy as np
import pytest

from jeicyboodsp_b200 import synth
from oracle.oracle import DenoiseParams, MfccParams

pytestmark = pytest.mark.skipif(
    not __import__("oracle.oracle", fromlist=["RefPrograms"]).RefPrograms().available(),
    reason="oracle/_ref not built (no reference checkout here)")


@pytest.mark.parametrize("n", [512, 1024])
def test_roundtrip_program_bit_exact(oracle, refprog, n):
    x = synth.roundtrip_signal(30_000 + 17)
    got, _ = oracle.roundtrip(x, n)
    ref = refprog.roundtrip(x, n)
    assert len(ref) == -(-len(x) // n) * n          # stale-tail rule: a full last block is written
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n", [256, 512, 2048, 32768])
@pytest.mark.parametrize("forward", [True, False])
def test_fftprocess_bit_exact(oracle, refprog, n, forward):
    rng = np.random.default_rng(n)
    z = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    assert np.array_equal(oracle.fftprocess(z, forward), refprog.fftprocess(z, forward))


def test_bitrev_and_dft(oracle, refprog):
    for n in (256, 512, 1024, 32768):
        assert np.array_equal(oracle.bitrev_table(n), refprog.bitrev_table(n).astype(np.int32))
    x = synth.roundtrip_signal(512)
    assert np.array_equal(oracle.dftprocess(x), refprog.dftprocess(x))


@pytest.mark.parametrize("preset", ["ref", "bench"])
@pytest.mark.parametrize("mode", [0, 1])
def test_denoise_programs_bit_exact(oracle, refprog, preset, mode):
    x = synth.denoise_stream(11, 64_000 + 123)
    res = oracle.denoise(x, DenoiseParams.preset(preset, mode))
    ref, energy, zcr = refprog.denoise(x, preset, mode, want_vad=True)
    assert len(res.publish) > 0, "noise path never fired: parity run would be vacuous"
    assert np.array_equal(res.out, ref)
    assert np.allclose(res.energy, energy, atol=1e-5)
    # the reference reads one element past its buffer (appendix C-3): its count is ours + {0, 1}
    assert set(np.unique(zcr - res.zcr)) <= {0, 1}


def test_fastconv_programs_bit_exact(oracle, refprog):
    x = synth.fastconv_source(7, 30_000)
    h = synth.hrir_pair(7)
    for ear in range(2):
        got, _ = oracle.fastconv(x, h[ear], 512, 1, 1024)
        assert np.array_equal(got, refprog.fastconv(x, "bench", np.concatenate([h[ear], [0.0]])))
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "fastconv.npz"))
    taps = np.zeros(7169)
    taps[g["ref_taps_idx"]] = g["ref_taps_val"]
    xr = synth.fastconv_source(1, 1024 * 11 + 5)
    got, _ = oracle.fastconv(xr, taps, 1024, 7, 8192)
    assert np.array_equal(got, refprog.fastconv(xr, "ref"))


@pytest.mark.parametrize("preset,ncep", [("ref", 12), ("mid", 13)])
def test_mfcc_programs(oracle, refprog, preset, ncep):
    x = synth.mfcc_utterance(3, 40_000 + 9)
    got = oracle.mfcc_program(x, MfccParams.preset(preset))
    ref = refprog.mfcc(x, preset, ncep)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-9


def test_pitch_program(oracle, refprog):
    """PitchEstimation_method1: arg bit-exact, dMax to the program's printf precision; the exact-integer
    scan (the GPU path's contract) agrees wherever the two best lags are further apart than the double FFT's rounding noise."""
    for x in (synth.denoise_stream(5, 50_000 + 123), synth.roundtrip_signal(20_000 + 5),
              np.random.default_rng(9).normal(0, 4000, 30_000).astype(np.int16), np.zeros(2_000, np.int16)):
        arg, mx = refprog.pitch(x)
        oa, om = oracle.pitch(x)
        assert len(arg) == -(-len(x) // 512)
        assert np.array_equal(arg, oa)
        assert np.allclose(mx, om, rtol=0, atol=1e-6 + 1e-15 * np.abs(om).max())
        ea, em = oracle.pitch(x, exact=True)
        assert np.array_equal(ea, oa)
        assert np.abs(em - om).max() <= 1e-9 * max(1.0, np.abs(em).max())


def test_mvdr_program(oracle, refprog):
    """BeamForming_MVDR_ver1: int16 output bit-exact against the unmodified program built over
    oracle/eigen_shim; also a stream whose matrix never leaves zero (all voice): the program's NaN -> (short) 0."""
    for s in (5, 6):
        xl, xr = synth.mvdr_pair(s, 40_000 + 123)
        out, pre, corr, vad = oracle.mvdr(xl, xr)
        assert (vad == 0).sum() > 2 and corr[-1, 0] > 0
        assert np.array_equal(out, refprog.mvdr(xl, xr))
        # the off-diagonal sums the closed form drops are rounding noise in the program too
        assert np.abs(corr[:, 1:3]).max() <= 1e-12 * corr[-1, 0]
    loud = np.random.default_rng(1).normal(0, 3000, 5_000).astype(np.int16)
    out, pre, corr, vad = oracle.mvdr(loud, loud[::-1].copy())
    ref = refprog.mvdr(loud, loud[::-1].copy())
    assert vad.all() and np.array_equal(out, ref) and not ref.any()
