"""TEST INFRASTRUCTURE ONLY -- ctypes doorway to the parity checker.

Two things live behind this module, neither of which the product ever imports:

* ``Oracle``: oracle/_build/libjdsp_oracle.so, the plain-C restatement (oracle/jdsp_oracle.c).
* ``RefPrograms``: the UNMODIFIED reference programs compiled by oracle/build.sh into oracle/_ref/
  (present in the build container and shipped to the GPU box as binaries; absent -> ``available()``
  is False and callers fall back to the committed fixtures in tests/golden/).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libjdsp_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

PI_FFT = 3.14159265358  # FFTAlgorithm_ver2.cpp:15
PI_DSP = 3.141592       # SpectralSubtraction_final.cpp:52 and every other program


def build(force: bool = False) -> None:
    """Compile the restatement (and oracle/_ref when the reference checkout is present)."""
    have_ref = os.path.isdir(os.environ.get("JDSP_REFERENCE_DIR", "/root/reference"))
    need = force or not os.path.exists(LIB_PATH) or (
        os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "jdsp_oracle.c")))
    if have_ref and not os.path.exists(os.path.join(REF_DIR, "ss_bench")):
        need = True
    if need:
        subprocess.run(["bash", os.path.join(HERE, "build.sh")], check=True, stdout=subprocess.DEVNULL)


class DenoiseParams(C.Structure):
    _fields_ = [("nfft", C.c_int32), ("hop", C.c_int32), ("mode", C.c_int32), ("zcr_thr", C.c_int32),
                ("noise_frames", C.c_int32), ("reserved", C.c_int32), ("win_a0", C.c_double),
                ("win_a1", C.c_double), ("pi", C.c_double), ("energy_thr", C.c_double)]

    @classmethod
    def preset(cls, name: str, mode: int) -> "DenoiseParams":
        if name == "ref":      # SpectralSubtraction_final.cpp:48-56,226
            return cls(1024, 512, mode, 200, 10, 0, 0.54, 0.46, PI_DSP, 700.0)
        if name == "bench":    # BASELINE.json config 2 (ZCR threshold rescaled to hop/4)
            return cls(512, 256, mode, 64, 10, 0, 0.5, 0.5, PI_DSP, 700.0)
        raise ValueError(name)


class MfccParams(C.Structure):
    _fields_ = [("frame_len", C.c_int32), ("hop", C.c_int32), ("nfft", C.c_int32), ("n_mel", C.c_int32),
                ("n_cep", C.c_int32), ("lifter", C.c_int32), ("half_sr", C.c_double), ("preemph", C.c_double),
                ("win_a0", C.c_double), ("win_a1", C.c_double), ("pi", C.c_double)]

    @classmethod
    def preset(cls, name: str) -> "MfccParams":
        if name == "ref":      # MFCCFeatureExtraction_auto_version1.cpp:23-33
            return cls(1024, 512, 1024, 38, 12, 22, 22050.0, 0.96, 0.54, 0.46, PI_DSP)
        if name == "mid":      # sed-built through the reference's own code
            return cls(512, 256, 512, 26, 13, 22, 8000.0, 0.96, 0.54, 0.46, PI_DSP)
        if name == "bench":    # BASELINE.json config 4
            return cls(400, 160, 512, 26, 13, 22, 8000.0, 0.96, 0.54, 0.46, PI_DSP)
        raise ValueError(name)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@dataclass
class DenoiseResult:
    out: np.ndarray          # int16 [(nb-2)*hop]
    out_f64: np.ndarray      # pre-cast doubles
    vad: np.ndarray          # uint8 [nb]
    zcr: np.ndarray          # int32 [nb]
    energy: np.ndarray       # float64 [nb]
    publish: np.ndarray      # int32 [n_publish] block indices


class Oracle:
    def __init__(self) -> None:
        build()
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.jo_roundtrip_i16.restype = C.c_long
        L.jo_denoise_i16.restype = C.c_long
        L.jo_fastconv_i16.restype = C.c_long
        L.jo_mfcc_frames.restype = C.c_long
        L.jo_mfcc_program.restype = C.c_long
        L.jo_pitch_i16.restype = C.c_long
        L.jo_pitch_exact_i16.restype = C.c_long
        L.jo_mvdr_i16.restype = C.c_long

    # ---- FFT ------------------------------------------------------------------------------------
    def bitrev_table(self, n: int) -> np.ndarray:
        t = np.zeros(n, np.int32)
        self.lib.jo_bitrev_table(C.c_int(n), _p(t, C.c_int32))
        return t

    def fftprocess(self, x: np.ndarray, forward: bool, pi: float = PI_FFT) -> np.ndarray:
        x = np.ascontiguousarray(x, np.complex128)
        out = np.empty_like(x)
        flat_in, flat_out = x.reshape(-1, x.shape[-1]), out.reshape(-1, x.shape[-1])
        n = x.shape[-1]
        for r in range(flat_in.shape[0]):
            self.lib.jo_fftprocess(_p(flat_in[r].view(np.float64), C.c_double),
                                   _p(flat_out[r].view(np.float64), C.c_double),
                                   C.c_int(n), C.c_int(1 if forward else 0), C.c_double(pi))
        return out

    def dft_exact(self, x: np.ndarray, sign: int) -> np.ndarray:
        x = np.ascontiguousarray(x, np.complex128)
        out = np.empty_like(x)
        self.lib.jo_dft_exact(_p(x.view(np.float64), C.c_double), _p(out.view(np.float64), C.c_double),
                              C.c_int(x.shape[-1]), C.c_int(sign))
        return out

    def dftprocess(self, x: np.ndarray, pi: float = PI_FFT) -> np.ndarray:
        x = np.ascontiguousarray(x, np.int16)
        out = np.empty(x.shape[-1], np.complex128)
        self.lib.jo_dftprocess(_p(x, C.c_int16), _p(out.view(np.float64), C.c_double), C.c_int(x.shape[-1]),
                               C.c_double(pi))
        return out

    # ---- programs -------------------------------------------------------------------------------
    def roundtrip(self, x: np.ndarray, nfft: int, pi: float = PI_FFT):
        x = np.ascontiguousarray(x, np.int16)
        nb = -(-len(x) // nfft)
        out = np.zeros(nb * nfft, np.int16)
        f64 = np.zeros(nb * nfft, np.float64)
        w = self.lib.jo_roundtrip_i16(_p(x, C.c_int16), C.c_long(len(x)), C.c_int(nfft), C.c_double(pi),
                                      _p(out, C.c_int16), _p(f64, C.c_double))
        return out[:w], f64[:w]

    def denoise(self, x: np.ndarray, params: DenoiseParams) -> DenoiseResult:
        x = np.ascontiguousarray(x, np.int16)
        nb = -(-len(x) // params.hop)
        n_out = max(nb - 2, 0) * params.hop
        out = np.zeros(n_out, np.int16)
        f64 = np.zeros(n_out, np.float64)
        vad = np.zeros(nb, np.uint8)
        zcr = np.zeros(nb, np.int32)
        en = np.zeros(nb, np.float64)
        pub = np.zeros(max(nb, 1), np.int32)
        npub = C.c_int32(0)
        w = self.lib.jo_denoise_i16(C.byref(params), _p(x, C.c_int16), C.c_long(len(x)), _p(out, C.c_int16),
                                    _p(f64, C.c_double), _p(vad, C.c_uint8), _p(zcr, C.c_int32),
                                    _p(en, C.c_double), _p(pub, C.c_int32), C.c_int32(len(pub)), C.byref(npub))
        assert w == n_out, (w, n_out)
        return DenoiseResult(out, f64, vad, zcr, en, pub[: npub.value].copy())

    def fastconv(self, x: np.ndarray, taps: np.ndarray, blk: int, q: int, nfft: int, ntaps: int | None = None):
        x = np.ascontiguousarray(x, np.int16)
        taps = np.ascontiguousarray(taps, np.float64)
        ntaps = q * blk + 1 if ntaps is None else ntaps
        full = np.zeros(ntaps, np.float64)
        full[: min(len(taps), ntaps)] = taps[:ntaps]
        nb = -(-len(x) // blk)
        n_out = max(nb - q, 0) * blk
        out = np.zeros(n_out, np.int16)
        f64 = np.zeros(n_out, np.float64)
        w = self.lib.jo_fastconv_i16(_p(x, C.c_int16), C.c_long(len(x)), C.c_int(blk), C.c_int(q), C.c_int(nfft),
                                     _p(full, C.c_double), C.c_int(ntaps), _p(out, C.c_int16), _p(f64, C.c_double))
        assert w == n_out, (w, n_out)
        return out, f64

    def mel_init(self, params: MfccParams):
        nbin = params.nfft // 2
        weight = np.zeros(nbin, np.float64)
        chan = np.zeros(nbin, np.int32)
        edges = np.zeros(params.n_mel + 1, np.float64)
        self.lib.jo_mel_init(C.byref(params), _p(weight, C.c_double), _p(chan, C.c_int32), _p(edges, C.c_double))
        return weight, chan, edges

    def mfcc_frames(self, s: np.ndarray, params: MfccParams) -> np.ndarray:
        s = np.ascontiguousarray(s, np.int16)
        nf = (len(s) - params.frame_len) // params.hop + 1 if len(s) >= params.frame_len else 0
        feat = np.zeros((max(nf, 0), params.n_cep), np.float64)
        got = self.lib.jo_mfcc_frames(C.byref(params), _p(s, C.c_int16), C.c_long(len(s)), _p(feat, C.c_double))
        assert got == nf, (got, nf)
        return feat

    # ---- pitch -------------------------------------------------------------------
    def pitch(self, x: np.ndarray, blk: int = 512, nfft: int = 1024, min_lag: int = 100, exact: bool = False):
        """(arg, rmax) per block; exact=False follows the reference's double-FFT arithmetic, exact=True scans the
        exact integer autocorrelation (the GPU path's contract)."""
        x = np.ascontiguousarray(x, np.int16)
        nb = -(-len(x) // blk)
        arg, mx = np.zeros(nb, np.int32), np.zeros(nb, np.float64)
        fn = self.lib.jo_pitch_exact_i16 if exact else self.lib.jo_pitch_i16
        got = fn(_p(x, C.c_int16), C.c_long(len(x)), C.c_int(blk), C.c_int(nfft), C.c_int(min_lag),
                 _p(arg, C.c_int32), _p(mx, C.c_double))
        assert got == nb, (got, nb)
        return arg, mx

    # ---- MVDR ---------------------------------------------------------------------
    def mvdr(self, xl: np.ndarray, xr: np.ndarray, dtime: float = 0.0):
        """(out int16, pre-cast double (NaN while the matrix is singular), corr [nb,4] spatial matrix in force for each block,
        vad [nb]) of BeamForming_MVDR_ver1."""
        xl, xr = np.ascontiguousarray(xl, np.int16), np.ascontiguousarray(xr, np.int16)
        n = min(len(xl), len(xr))
        nb = -(-n // 512)
        out = np.zeros(max(nb - 1, 0) * 512 + 8, np.int16)
        pre = np.zeros(max(nb - 1, 0) * 512 + 8, np.float64)
        corr, vad = np.zeros((max(nb, 1), 4), np.float64), np.zeros(max(nb, 1), np.uint8)
        w = self.lib.jo_mvdr_i16(_p(xl, C.c_int16), _p(xr, C.c_int16), C.c_long(n), C.c_double(dtime), _p(out, C.c_int16),
                                 _p(pre, C.c_double), _p(corr, C.c_double), _p(vad, C.c_uint8))
        assert w == max(nb - 1, 0) * 512, (w, nb)
        return out[:w], pre[:w], corr[:nb], vad[:nb]

    def mfcc_program(self, x: np.ndarray, params: MfccParams) -> np.ndarray:
        x = np.ascontiguousarray(x, np.int16)
        nb = -(-len(x) // (2 * params.hop))
        feat = np.zeros((max(2 * nb - 1, 0), params.n_cep), np.float64)
        rows = self.lib.jo_mfcc_program(C.byref(params), _p(x, C.c_int16), C.c_long(len(x)), _p(feat, C.c_double))
        assert rows == feat.shape[0], (rows, feat.shape)
        return feat


class RefPrograms:
    """Runs the compiled, unmodified reference programs (oracle/_ref) on in-memory arrays."""

    WAV_HEADER = bytes(44)  # FFT / fast-conv / MFCC skip 44 bytes (e.g. FFTAlgorithm_ver2.cpp:59); content unused

    def __init__(self, ref_dir: str = REF_DIR) -> None:
        self.dir = ref_dir

    def available(self, name: str = "ss_bench") -> bool:
        return os.path.exists(os.path.join(self.dir, name))

    def _run(self, exe: str, args: list[str], capture: bool = False) -> str:
        # every main ends in getchar() -> stdin from /dev/null; per-frame printf -> discarded unless wanted
        r = subprocess.run([os.path.join(self.dir, exe)] + args, stdin=subprocess.DEVNULL,
                           stdout=subprocess.PIPE if capture else subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                           check=True)
        return r.stdout.decode("latin1") if capture else ""

    def roundtrip(self, x: np.ndarray, nfft: int) -> np.ndarray:
        with tempfile.TemporaryDirectory() as d:
            fi, fo = os.path.join(d, "in.wav"), os.path.join(d, "out.pcm")
            with open(fi, "wb") as f:
                f.write(self.WAV_HEADER + np.ascontiguousarray(x, np.int16).tobytes())
            self._run(f"fft_roundtrip_{nfft}", [fi, fo])
            return np.fromfile(fo, np.int16)

    def denoise(self, x: np.ndarray, preset: str, mode: int, want_vad: bool = False):
        exe = ("ss" if mode == 0 else "wiener") + "_" + preset
        with tempfile.TemporaryDirectory() as d:
            fi, fo = os.path.join(d, "in.pcm"), os.path.join(d, "out.pcm")
            np.ascontiguousarray(x, np.int16).tofile(fi)
            log = self._run(exe, [fi, fo], capture=want_vad)
            out = np.fromfile(fo, np.int16)
        if not want_vad:
            return out
        en, zc = [], []
        for line in log.splitlines():       # " dEnergy %f , dZCR %d" (SpectralSubtraction_final.cpp:146)
            if "dEnergy" in line:
                parts = line.replace(",", " ").split()
                en.append(float(parts[1]))
                zc.append(int(parts[3]))
        return out, np.array(en), np.array(zc, np.int32)

    def fastconv(self, x: np.ndarray, preset: str, taps: np.ndarray | None = None) -> np.ndarray:
        with tempfile.TemporaryDirectory() as d:
            fi, fo, ft = os.path.join(d, "in.wav"), os.path.join(d, "out.pcm"), os.path.join(d, "taps.f64")
            with open(fi, "wb") as f:
                f.write(self.WAV_HEADER + np.ascontiguousarray(x, np.int16).tobytes())
            args = [fi, fo]
            if taps is not None:
                np.ascontiguousarray(taps, np.float64).tofile(ft)
                args.append(ft)
            self._run(f"fastconv_{preset}", args)
            return np.fromfile(fo, np.int16)

    def mfcc(self, x: np.ndarray, preset: str, n_cep: int) -> np.ndarray:
        with tempfile.TemporaryDirectory() as d:
            fi, fo, fl = os.path.join(d, "in.wav"), os.path.join(d, "out.mfc"), os.path.join(d, "list.txt")
            with open(fi, "wb") as f:
                f.write(self.WAV_HEADER + np.ascontiguousarray(x, np.int16).tobytes())
            with open(fl, "w") as f:
                f.write(f"{fi} {fo}")   # NO trailing newline (feof loop would run again)
            self._run(f"mfcc_{preset}", [fl])
            return np.fromfile(fo, np.float64).reshape(-1, n_cep)

    def pitch(self, x: np.ndarray):
        """PitchEstimation_method1: parse 'Estimation arg %d , dMin %f pitch %f' (:109), one line per block."""
        with tempfile.TemporaryDirectory() as d:
            fi = os.path.join(d, "in.wav")
            with open(fi, "wb") as f:
                f.write(self.WAV_HEADER + np.ascontiguousarray(x, np.int16).tobytes())
            log = self._run("pitch_ref", [fi], capture=True)
        arg, mx = [], []
        for line in log.splitlines():
            if line.startswith("Estimation arg"):
                parts = line.replace(",", " ").split()
                arg.append(int(parts[2]))
                mx.append(float(parts[4]))
        return np.array(arg, np.int32), np.array(mx, np.float64)

    def mvdr(self, xl: np.ndarray, xr: np.ndarray) -> np.ndarray:
        """BeamForming_MVDR_ver1 <left.wav> <right.wav> <out.pcm> (angle 0 is hard-wired in the program, :58-60)."""
        with tempfile.TemporaryDirectory() as d:
            fl, fr, fo = os.path.join(d, "l.wav"), os.path.join(d, "r.wav"), os.path.join(d, "out.pcm")
            for fn, x in ((fl, xl), (fr, xr)):
                with open(fn, "wb") as f:
                    f.write(self.WAV_HEADER + np.ascontiguousarray(x, np.int16).tobytes())
            self._run("mvdr_ref", [fl, fr, fo])
            return np.fromfile(fo, np.int16)

    def fftprocess(self, x: np.ndarray, forward: bool) -> np.ndarray:
        """The reference's own FFTProcess built with BLOCK_LEN == len(x) (valid for 2^8..2^15)."""
        x = np.ascontiguousarray(x, np.complex128)
        n = x.shape[-1]
        lib = C.CDLL(os.path.join(self.dir, f"libfftprocess_{n}.so"))
        assert lib.jref_block_len() == n
        out = np.zeros_like(x)
        fi, fo = x.reshape(-1, n), out.reshape(-1, n)
        for r in range(fi.shape[0]):
            lib.jref_fftprocess(_p(fi[r].view(np.float64), C.c_double), _p(fo[r].view(np.float64), C.c_double),
                                C.c_int(n), C.c_int(1 if forward else 0))
        return out

    def dftprocess(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.int16)
        n = x.shape[-1]
        lib = C.CDLL(os.path.join(self.dir, f"libfftprocess_{n}.so"))
        out = np.zeros(n, np.complex128)
        lib.jref_dftprocess(_p(x, C.c_int16), _p(out.view(np.float64), C.c_double), C.c_int(n))
        return out

    def bitrev_table(self, n: int) -> np.ndarray:
        lib = C.CDLL(os.path.join(self.dir, f"libfftprocess_{n}.so"))
        t = np.zeros(n, np.int16)
        lib.jref_bitrev_table(_p(t, C.c_int16), C.c_int(n))
        return t
