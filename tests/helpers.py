"""Parity rules of SURVEY.md section 8c, in one place."""
from __future__ import annotations

import numpy as np

TOL_REL_PEAK = 1e-4   # north_star: max abs error <= 1e-4 relative to signal peak
MIN_SNR_DB = 90.0     # ... or >= 90 dB SNR against the reference output


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    err = np.asarray(got, np.float64) - ref
    den = float(np.sum(err * err))
    return float("inf") if den == 0 else 10.0 * np.log10(float(np.sum(ref * ref)) / den)


def assert_float_parity(got, ref, what: str = "") -> None:
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    peak = float(np.abs(ref).max()) or 1.0
    err = float(np.abs(got - ref).max())
    assert err <= TOL_REL_PEAK * peak or snr_db(ref, got) >= MIN_SNR_DB, (what, err, peak, snr_db(ref, got))


def assert_i16_parity(got: np.ndarray, ref: np.ndarray, max_flip_frac: float, what: str = "") -> None:
    """(short) casts truncate, so a float pipeline may land one LSB away at truncation boundaries
    (SURVEY 0.3-1).  Never more than one LSB, never a wrap, and only on a small fraction of samples."""
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max(initial=0) <= 1, (what, int(d.max()), int(np.argmax(d)))
    frac = float((d > 0).mean()) if d.size else 0.0
    assert frac <= max_flip_frac, (what, frac)
