"""The C++ drop-in console programs (programs/, include/jdsp_dropin.hpp) against the reference fixtures:
same argv conventions and file formats as the reference programs, GPU routine behind the C ABI."""
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_float_parity, assert_i16_parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "programs", "_build")
HDR = bytes(44)


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "programs")], check=True, stdout=subprocess.DEVNULL)
    return BIN


def test_programs_build_without_cuda_headers(built):
    for p in ("jdsp_fft_roundtrip", "jdsp_denoise", "jdsp_fastconv", "jdsp_mfcc", "jdsp_blockwise", "jdsp_pitch", "jdsp_mvdr"):
        assert os.path.exists(os.path.join(built, p))


def _run(built, prog, *args):
    subprocess.run([os.path.join(built, prog)] + list(args), check=True, stdin=subprocess.DEVNULL)


@pytest.mark.gpu
def test_roundtrip_program(built, tmp_path):
    g = np.load(os.path.join(G, "fft.npz"))
    fi, fo = tmp_path / "in.wav", tmp_path / "out.pcm"
    fi.write_bytes(HDR + g["pcm"].tobytes())
    for n in (512, 1024):
        _run(built, "jdsp_fft_roundtrip", str(fi), str(fo), str(n))
        assert_i16_parity(np.fromfile(fo, np.int16), g[f"rt{n}"], max_flip_frac=0.6)


@pytest.mark.gpu
@pytest.mark.parametrize("prog", ["jdsp_denoise", "jdsp_blockwise"])
@pytest.mark.parametrize("preset", ["ref", "bench"])
def test_denoise_programs(built, tmp_path, prog, preset):
    g = np.load(os.path.join(G, "denoise.npz"))
    fi, fo = tmp_path / "in.pcm", tmp_path / "out.pcm"
    x = g["pcm_3"] if prog == "jdsp_denoise" else g["pcm_3"][:20_000 + 77]   # block-at-a-time path: keep it short
    x.tofile(fi)
    for nm in ("ss", "wiener"):
        _run(built, prog, nm, preset, str(fi), str(fo))
        got = np.fromfile(fo, np.int16)
        ref = g[f"{nm}_{preset}_3"][: len(got)]
        if prog == "jdsp_blockwise":   # the golden was made from the longer file: compare the common whole blocks
            hop = 512 if preset == "ref" else 256
            keep = (len(x) // hop - 2) * hop
            got, ref = got[:keep], ref[:keep]
        assert_i16_parity(got, ref, max_flip_frac=2e-3, what=f"{prog} {nm} {preset}")


@pytest.mark.gpu
def test_fastconv_program(built, tmp_path):
    g = np.load(os.path.join(G, "fastconv.npz"))
    fi, fo, ft = tmp_path / "in.wav", tmp_path / "out.pcm", tmp_path / "taps.f64"
    fi.write_bytes(HDR + g["pcm_bench"].tobytes())
    np.concatenate([g["hrir_bench"], np.zeros((2, 1))], axis=1).tofile(ft)
    _run(built, "jdsp_fastconv", "bench", str(fi), str(fo), str(ft))
    assert_i16_parity(np.fromfile(fo, np.int16), g["out_bench_ear0"], max_flip_frac=2e-3)
    assert_i16_parity(np.fromfile(str(fo) + ".ear1", np.int16), g["out_bench_ear1"], max_flip_frac=2e-3)
    fi.write_bytes(HDR + g["pcm_ref"].tobytes())
    taps = np.zeros(7169); taps[g["ref_taps_idx"]] = g["ref_taps_val"]
    taps.tofile(ft)
    _run(built, "jdsp_fastconv", "ref", str(fi), str(fo), str(ft))
    assert_i16_parity(np.fromfile(fo, np.int16), g["out_ref"], max_flip_frac=1e-2)


@pytest.mark.gpu
def test_mfcc_program_writes_the_mfc_format(built, tmp_path):
    g = np.load(os.path.join(G, "mfcc.npz"))
    fi, fo, fl = tmp_path / "in.wav", tmp_path / "out.mfc", tmp_path / "list.txt"
    fi.write_bytes(HDR + g["pcm"].tobytes())
    fl.write_text(f"{fi} {fo}\n")          # a trailing newline is fine here (it crashes the reference)
    for preset, ncep in (("ref", 12), ("mid", 13)):
        _run(built, "jdsp_mfcc", str(fl), preset)
        rows = np.fromfile(fo, np.float64).reshape(-1, ncep)   # raw double[n_cep] rows: what GMMAlgorithm_* read
        assert_float_parity(rows, g[preset], f"mfcc program {preset}")


@pytest.mark.gpu
@pytest.mark.parametrize("how", ["batched", "block"])
def test_pitch_program_prints_the_reference_lines(built, tmp_path, how):
    """Same stdout lines as PitchEstimation_method1 (:109): arg bit-exact, dMax / pitch as printed."""
    g = np.load(os.path.join(G, "pitch.npz"))
    fi = tmp_path / "in.wav"
    x = g["pcm_3"] if how == "batched" else g["pcm_3"][:6 * 512 + 100]
    fi.write_bytes(HDR + x.tobytes())
    args = [os.path.join(built, "jdsp_pitch"), str(fi)] + (["block"] if how == "block" else [])
    out = subprocess.run(args, check=True, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE).stdout.decode()
    rows = [ln.replace(",", " ").split() for ln in out.splitlines() if ln.startswith("Estimation arg")]
    arg = np.array([int(r[2]) for r in rows])
    mx = np.array([float(r[4]) for r in rows])
    pitch = np.array([float(r[6]) for r in rows])
    nb = -(-len(x) // 512)
    assert len(arg) == nb and out.rstrip().endswith("Processing End")
    if how == "batched":
        assert np.array_equal(arg, g["arg_3"])
        assert np.allclose(mx, g["rmax_3"], rtol=1e-12, atol=1e-5)
    else:   # a shorter file: whole blocks agree with the fixture, the short last block is checked for the stale-tail rule
        assert np.array_equal(arg[: nb - 1], g["arg_3"][: nb - 1])
    assert np.allclose(pitch, 16000.0 / arg, rtol=0, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("how", ["batched", "block"])
def test_mvdr_program(built, tmp_path, how):
    """Same argv and file formats as BeamForming_MVDR_ver1; the block-at-a-time form goes through MvdrStream::ProcessMVDR."""
    g = np.load(os.path.join(G, "mvdr.npz"))
    fl, fr, fo = tmp_path / "l.wav", tmp_path / "r.wav", tmp_path / "out.pcm"
    n = len(g["left_3"]) if how == "batched" else 12 * 512 + 100
    fl.write_bytes(HDR + g["left_3"][:n].tobytes())
    fr.write_bytes(HDR + g["right_3"][:n].tobytes())
    _run(built, "jdsp_mvdr", str(fl), str(fr), str(fo), *(["block"] if how == "block" else []))
    got = np.fromfile(fo, np.int16)
    nb = -(-n // 512)
    assert len(got) == (nb - 1) * 512
    keep = len(got) if how == "batched" else (nb - 2) * 512     # the shorter file's last block has its own stale tail
    assert got[:keep].any()
    assert_i16_parity(got[:keep], g["out_3"][:keep], max_flip_frac=2e-2, what=f"mvdr program {how}")
