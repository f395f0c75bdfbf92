"""Known-answer tests the reference lacks, run on the oracle."""
import numpy as np
import pytest

from jeicyboodsp_b200 import synth
from oracle.oracle import PI_FFT, DenoiseParams, MfccParams


@pytest.mark.parametrize("n", [8, 256, 1024])
def test_fft_impulse_dc_tone(oracle, n):
    imp = np.zeros(n, complex); imp[0] = 1
    assert np.allclose(oracle.fftprocess(imp, True), np.ones(n), atol=1e-9)
    dc = np.ones(n, complex)
    X = oracle.fftprocess(dc, True)
    assert abs(X[0] - n) < 1e-7 and np.abs(X[1:]).max() < 1e-6
    k = 3 % n
    tone = np.exp(2j * np.pi * k * np.arange(n) / n)
    X = oracle.fftprocess(tone, True)
    assert abs(X[k] - n) < 1e-6 and np.abs(np.delete(X, k)).max() < 1e-6


def test_fftprocess_equals_dftprocess_and_parseval(oracle):
    x = synth.roundtrip_signal(512)
    X = oracle.fftprocess(x.astype(complex), True)
    # the author's own cross-check (:74); DFTProcess feeds angles up to 2*PI*511*511/512 to cos/sin, so the
    # truncated PI literal shows up ~250x larger there (5e-9) than in FFTProcess (2e-11)
    assert np.abs(X - oracle.dftprocess(x)).max() / np.abs(X).max() < 5e-8
    assert abs(np.sum(np.abs(X) ** 2) / 512 - np.sum(x.astype(float) ** 2)) / np.sum(x.astype(float) ** 2) < 1e-9
    # the PI literal (3.14159265358) leaves a 2-5e-11 relative deviation from an exact DFT
    dev = np.abs(X - np.fft.fft(x)).max() / np.abs(X).max()
    assert 1e-13 < dev < 1e-9
    assert PI_FFT == 3.14159265358


def test_unnormalised_inverse(oracle):
    z = np.random.default_rng(0).normal(size=256) + 0j
    back = oracle.fftprocess(oracle.fftprocess(z, True), False)
    assert np.allclose(back, 256 * z, atol=1e-6)


def test_block_counts_and_warmup(oracle):
    x = synth.roundtrip_signal(160_000)
    assert len(oracle.roundtrip(x, 512)[0]) == 160_256
    r = oracle.denoise(x, DenoiseParams.preset("ref", 0))
    assert len(r.out) == 159_232                                       # (nb-2)*H, SURVEY A.2
    assert len(oracle.denoise(x, DenoiseParams.preset("bench", 0)).out) == 159_488
    assert oracle.mfcc_program(x, MfccParams.preset("ref")).shape == (313, 12)


@pytest.mark.parametrize("preset,gain", [("ref", 1.08), ("bench", 1.0)])
def test_ola_identity_without_noise_estimate(oracle, preset, gain):
    """A loud tone is always 'voice', so the noise spectrum stays 0 and the filter is the identity: the
    output is the input delayed by one hop times the window overlap sum (1.08 for Hamming, 1.0 for Hann)."""
    p = DenoiseParams.preset(preset, 0)
    n = 20 * p.hop
    x = np.round(9000 * np.sin(2 * np.pi * 440 * np.arange(n) / 16000)).astype(np.int16)
    r = oracle.denoise(x, p)
    assert len(r.publish) == 0 and r.vad.all()
    i = np.arange(p.nfft)
    w = p.win_a0 - p.win_a1 * np.cos(2 * p.pi * i / (p.nfft - 1))
    m = np.arange(len(r.out_f64))
    expect = x[m + p.hop] * (w[p.hop + (m % p.hop)] + w[m % p.hop])
    assert np.abs(r.out_f64 - expect).max() < 1e-6
    assert abs(np.mean(w[: p.hop] + w[p.hop:]) - gain) < 0.01


def test_overlap_save_equals_direct_convolution(oracle):
    x = synth.fastconv_source(9, 20 * 512)
    h = synth.hrir_pair(9)[0]
    out, f64 = oracle.fastconv(x, h, 512, 1, 1024)
    xp = x.astype(float).copy(); xp[:512] = 0                         # first block never enters the history (C-4)
    full = np.convolve(xp, h)
    assert np.abs(f64 - full[512:512 + len(f64)]).max() < 1e-8


def test_mel_table_structure(oracle):
    for preset, first_edge in (("ref", 65.4), ("bench", None)):
        p = MfccParams.preset(preset)
        w, ch, edges = oracle.mel_init(p)
        assert np.all(np.diff(ch) >= 0) and np.all(np.diff(ch) <= 1) and ch.max() <= p.n_mel   # C-14
        assert np.all((w >= 0) & (w <= 1))
        assert abs(edges[-1] - p.half_sr) < 1e-6
        if first_edge:
            assert abs(edges[0] - first_edge) < 0.05
