"""The drop-in boundary: libjdsp.so loads, exports every symbol include/jdsp.h declares, and refuses to
compute without a CUDA device (no CPU fallback).  No compute calls here -- this runs without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "jdsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jdsp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from jeicyboodsp_b200 import build
    build.build()          # nvcc cross-compiles sm_100a without a GPU
    from jeicyboodsp_b200.binding import Library
    return Library()


def test_exports_every_declared_symbol(lib):
    from jeicyboodsp_b200.binding import ABI_SYMBOLS
    declared = _header_functions()
    assert len(declared) >= 40
    assert sorted(ABI_SYMBOLS) == declared, "binding.ABI_SYMBOLS must list exactly what the header declares"
    for name in declared:
        assert hasattr(lib.lib, name), f"libjdsp.so does not export {name}"
    assert lib.lib.jdsp_abi_version() == 5


def test_struct_layouts_match_header(lib):
    from jeicyboodsp_b200.binding import DenoiseParams, FastconvParams, MfccParams, MvdrParams, PitchParams
    assert C.sizeof(PitchParams) == 4 * 4 + 8
    assert C.sizeof(MvdrParams) == 4 * 4 + 6 * 8
    v = lib.mvdr_params("ref")         # BeamForming_MVDR_ver1.cpp:31-40,58-60
    assert (v.n_fft, v.block, v.keep, v.energy_thr, v.fs, v.dtime, v.win_a0, v.win_a1, v.pi_literal) == \
        (1024, 512, 511, 700.0, 16000.0, 0.0, 0.54, 0.46, 3.141592)
    assert C.sizeof(DenoiseParams) == 6 * 4 + 4 * 8
    assert C.sizeof(FastconvParams) == 6 * 4
    assert C.sizeof(MfccParams) == 6 * 4 + 5 * 8
    p = lib.denoise_params("ref", 0)   # SpectralSubtraction_final.cpp:48-56
    assert (p.n_fft, p.hop, p.zcr_thr, p.noise_frames, p.win_a0, p.win_a1, p.pi_literal, p.energy_thr) == \
        (1024, 512, 200, 10, 0.54, 0.46, 3.141592, 700.0)
    p = lib.denoise_params("bench", 1)
    assert (p.n_fft, p.hop, p.zcr_thr, p.win_a0, p.win_a1, p.mode) == (512, 256, 64, 0.5, 0.5, 1)
    f = lib.fastconv_params("ref")     # Fast_Convolution_Based_3DAudio_Impl.cpp:47-48, FilterCoefficient.h:1-2
    assert (f.block, f.n_fft, f.history_blocks, f.n_taps, f.n_ears) == (1024, 8192, 7, 7169, 1)
    m = lib.mfcc_params("ref")         # MFCCFeatureExtraction_auto_version1.cpp:23-33
    assert (m.frame_len, m.hop, m.n_fft, m.n_mel, m.n_cep, m.lifter, m.half_sr) == (1024, 512, 1024, 38, 12, 22, 22050.0)
    m = lib.mfcc_params("bench")
    assert (m.frame_len, m.hop, m.n_fft, m.n_mel, m.n_cep, m.half_sr) == (400, 160, 512, 26, 13, 8000.0)


def test_bitrev_table_is_host_side_and_exact(lib, oracle):
    for n in (2, 512, 32768, 65536):
        assert (lib.bitrev_table(n) == oracle.bitrev_table(n)).all()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from jeicyboodsp_b200.binding import Context, JdspError
    with pytest.raises(JdspError) as e:
        Context(lib, 0)
    assert e.value.code == -3      # JDSP_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value) or "cudaGetDeviceCount" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under jeicyboodsp_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "jeicyboodsp_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in src.lower(), (dirpath, fn)
