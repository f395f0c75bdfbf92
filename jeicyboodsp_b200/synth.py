"""Synthetic workloads of BASELINE.json, numpy for CPU-sized cases and torch
for the device-resident bench tensors.  There is no dataset access; every input is generated."""
from __future__ import annotations

import numpy as np


# ---- config 1: FFT round trip ---------------------------------------------------------------------
def roundtrip_signal(n: int = 160_000, fs: float = 16_000.0, seed: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    x = 12000 * np.sin(2 * np.pi * 440 * t) + 6000 * np.sin(2 * np.pi * 1234.5 * t) + rng.normal(0, 1500, n)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


# ---- config 2: speech + AWGN streams ----------------------------------------------------------------
def _speech_envelope(t, xp):
    tp = xp.remainder(t, 2.0)
    env = 0.5 * (1.0 - xp.cos(2 * np.pi * (tp - 0.6) / 1.4))
    return xp.where(tp < 0.6, xp.zeros_like(env), env)


def denoise_stream(stream: int, n: int, fs: float = 16_000.0, sigma: float = 40.0, amp: float = 6000.0,
                   seed: int = 2) -> np.ndarray:
    """Gated 11-harmonic 'speech' + N(0, sigma).  The first 0.6 s of every 2 s is noise only so that
    runs of >= 10 non-voice blocks occur and the noise spectrum really gets published."""
    rng = np.random.default_rng([seed, stream])
    t = np.arange(n) / fs
    f0 = 120.0 + 30.0 * np.sin(2 * np.pi * 0.7 * t) + (stream % 40)
    phi = 2 * np.pi * np.cumsum(f0) / fs
    sp = np.zeros(n)
    for k in range(1, 12):
        sp += (amp / k) * np.sin(k * phi)
    x = 0.5 * sp * _speech_envelope(t, np) + rng.normal(0, sigma, n)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def denoise_streams_torch(n_streams: int, n: int, device, stream0: int = 0, fs: float = 16_000.0,
                          sigma: float = 40.0, amp: float = 6000.0, seed: int = 2, chunk: int = 256):
    """Device-resident [n_streams, n] int16 of the same family (torch generator, so the noise differs
    from the numpy version; parity streams are copied back to the host by the tests)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + stream0)
    out = torch.empty((n_streams, n), dtype=torch.int16, device=device)
    t = torch.arange(n, device=device, dtype=torch.float64) / fs
    env = _speech_envelope(t, torch).to(torch.float32)
    base = (120.0 + 30.0 * torch.sin(2 * np.pi * 0.7 * t))
    for s0 in range(0, n_streams, chunk):
        s1 = min(s0 + chunk, n_streams)
        ids = torch.arange(stream0 + s0, stream0 + s1, device=device)
        f0 = base[None, :] + (ids % 40).to(torch.float64)[:, None]
        phi = (2 * np.pi / fs) * torch.cumsum(f0, dim=1)
        phi = torch.remainder(phi, 2 * np.pi).to(torch.float32)
        sp = torch.zeros((s1 - s0, n), dtype=torch.float32, device=device)
        for k in range(1, 12):
            sp += (amp / k) * torch.sin(k * phi)
        x = 0.5 * sp * env[None, :]
        x += sigma * torch.randn((s1 - s0, n), generator=g, device=device, dtype=torch.float32)
        out[s0:s1] = torch.clamp(torch.round(x), -32768, 32767).to(torch.int16)
        del f0, phi, sp, x
    return out


# ---- config 3: sources + HRIR pairs -------------------------------------------------------------------
def fastconv_source(source: int, n: int, fs: float = 48_000.0, seed: int = 3) -> np.ndarray:
    rng = np.random.default_rng([seed, source])
    white = rng.normal(0, 1.0, n)
    # pink-ish: one-pole low-pass mixed with white, plus a tone
    lp = np.empty(n)
    acc = 0.0
    a = 0.98
    for i in range(n):  # CPU-sized inputs only
        acc = a * acc + (1 - a) * white[i]
        lp[i] = acc
    t = np.arange(n) / fs
    x = 2500.0 * (lp / (lp.std() + 1e-12)) * 0.6 + 800.0 * white * 0.4 + 2500.0 * np.sin(
        2 * np.pi * (300.0 + 7.0 * (source % 64)) * t)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def hrir_pair(source: int, taps: int = 512, seed: int = 3) -> np.ndarray:
    """[2, taps] float64: h[8]=1 then an exponentially decaying Gaussian tail (tau = 60 taps), each ear
    normalised to sum|h| <= 3 so the int16 output cannot wrap."""
    rng = np.random.default_rng([seed, 7919, source])
    h = np.zeros((2, taps))
    k = np.arange(taps)
    for ear in range(2):
        tail = rng.normal(0, 0.35, taps) * np.exp(-k / 60.0)
        tail[: 9 + ear] = 0.0          # small inter-aural delay
        h[ear] = tail
        h[ear, 8 + ear] = 1.0
        s = np.abs(h[ear]).sum()
        if s > 3.0:
            h[ear] *= 3.0 / s
    return h


# ---- config 4: utterances ---------------------------------------------------------------------------------
def mfcc_utterance(utt: int, n: int = 160_000, fs: float = 16_000.0, seed: int = 4) -> np.ndarray:
    """Same speech family as config 2 but with noise sigma >= 5 everywhere (no digital silence -> no ln 0)."""
    return denoise_stream(utt, n, fs=fs, sigma=25.0, amp=6000.0, seed=seed)


# ---- MVDR: two-microphone recordings -----------------------------------------------------
def mvdr_pair(stream: int, n: int, delay: int = 3, gain: float = 0.8, sigma_r: float = 35.0, seed: int = 6):
    """Left microphone = a config-2 stream (speech gated off for 0.6 s of every 2 s, so runs of non-voice blocks feed the
    spatial matrix); right microphone = the same scene `delay` samples later at `gain`, plus its own sensor noise."""
    left = denoise_stream(stream, n)
    rng = np.random.default_rng([seed, stream])
    late = np.concatenate([np.zeros(delay), left[:n - delay].astype(np.float64)]) if delay else left.astype(np.float64)
    right = gain * late + rng.normal(0, sigma_r, n)
    return left, np.clip(np.round(right), -32768, 32767).astype(np.int16)
