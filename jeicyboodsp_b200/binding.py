"""ctypes binding of the C ABI in include/jdsp.h (libjdsp.so, nvcc-built for sm_100a).

This is plumbing for tests and bench.py -- the product is the shared library.  There is no CPU
fallback: loading fails loudly when the library is missing, and every call fails with
JDSP_ERR_NO_DEVICE when no CUDA device is present.  (`Library(path=...)` lets tests/emul point the same
binding at the CPU execution emulator build used to debug kernels without a GPU.)

Pointer arguments accept numpy arrays (host memory) or torch tensors (host or CUDA); the caller is
responsible for passing device tensors to `_dev` entry points.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libjdsp.so")

SS, WIENER = 0, 1


class JdspError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libjdsp error {code}: {msg}")
        self.code = code


class DenoiseParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("hop", C.c_int32), ("mode", C.c_int32), ("zcr_thr", C.c_int32),
                ("noise_frames", C.c_int32), ("reserved", C.c_int32), ("win_a0", C.c_double),
                ("win_a1", C.c_double), ("pi_literal", C.c_double), ("energy_thr", C.c_double)]


class FastconvParams(C.Structure):
    _fields_ = [("block", C.c_int32), ("n_fft", C.c_int32), ("history_blocks", C.c_int32), ("n_taps", C.c_int32),
                ("n_ears", C.c_int32), ("shared_filter", C.c_int32)]


class MfccParams(C.Structure):
    _fields_ = [("frame_len", C.c_int32), ("hop", C.c_int32), ("n_fft", C.c_int32), ("n_mel", C.c_int32),
                ("n_cep", C.c_int32), ("lifter", C.c_int32), ("half_sr", C.c_double), ("preemph", C.c_double),
                ("win_a0", C.c_double), ("win_a1", C.c_double), ("pi_literal", C.c_double)]


class PitchParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("block", C.c_int32), ("min_lag", C.c_int32), ("reserved", C.c_int32),
                ("fs", C.c_double)]


class MvdrParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("block", C.c_int32), ("keep", C.c_int32), ("reserved", C.c_int32),
                ("energy_thr", C.c_double), ("fs", C.c_double), ("dtime", C.c_double), ("win_a0", C.c_double),
                ("win_a1", C.c_double), ("pi_literal", C.c_double)]


def _ptr(x):
    """Raw address of a numpy array / torch tensor / int / None."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"] or x.ndim <= 1 or x.strides[-1] == x.itemsize
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


# every symbol include/jdsp.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "jdsp_abi_version", "jdsp_last_error", "jdsp_device_count", "jdsp_create", "jdsp_create_on_stream",
    "jdsp_destroy", "jdsp_sync", "jdsp_cuda_stream", "jdsp_kernel_launches", "jdsp_malloc", "jdsp_free",
    "jdsp_host_alloc", "jdsp_host_free", "jdsp_memcpy_h2d", "jdsp_memcpy_d2h", "jdsp_fft_process",
    "jdsp_fft_c2c_f32", "jdsp_fft_c2c_f32_host", "jdsp_fft_c2c_f64", "jdsp_bitrev_table", "jdsp_roundtrip_i16_dev", "jdsp_roundtrip_i16",
    "jdsp_roundtrip_batch_i16", "jdsp_fastconv_i16_host", "jdsp_mfcc_frames_i16",
    "jdsp_denoise_params_preset", "jdsp_denoise_state_create", "jdsp_denoise_state_reset",
    "jdsp_denoise_state_destroy", "jdsp_denoise_i16_dev", "jdsp_denoise_i16", "jdsp_denoise_publish_counts",
    "jdsp_fastconv_params_preset", "jdsp_fastconv_state_create", "jdsp_fastconv_state_reset",
    "jdsp_fastconv_state_destroy", "jdsp_fastconv_i16_dev", "jdsp_fastconv_mix_i16_dev", "jdsp_fastconv_i16",
    "jdsp_mfcc_params_preset", "jdsp_mfcc_plan_create", "jdsp_mfcc_plan_destroy", "jdsp_mfcc_plan_tables",
    "jdsp_mfcc_frames_i16_dev", "jdsp_mfcc_frames_i16_scatter_dev", "jdsp_mfcc_frames_i16_multicast_dev", "jdsp_mfcc_program_i16",
    "jdsp_peer_export", "jdsp_peer_open", "jdsp_peer_close",
    "jdsp_pitch_params_preset", "jdsp_pitch_state_create", "jdsp_pitch_state_reset", "jdsp_pitch_state_destroy",
    "jdsp_pitch_i16_dev", "jdsp_pitch_i16",
    "jdsp_mvdr_params_preset", "jdsp_mvdr_state_create", "jdsp_mvdr_state_reset", "jdsp_mvdr_state_destroy",
    "jdsp_mvdr_i16_dev", "jdsp_mvdr_spatial_corr", "jdsp_mvdr_i16",
]


class Library:
    def __init__(self, path: str | None = None):
        self.path = path or DEFAULT_LIB
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                f"{self.path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        self.lib = C.CDLL(self.path)
        self.lib.jdsp_last_error.restype = C.c_char_p
        self.lib.jdsp_cuda_stream.restype = C.c_void_p

    def check(self, rc: int) -> None:
        if rc != 0:
            raise JdspError(rc, (self.lib.jdsp_last_error() or b"").decode())

    def device_count(self) -> int:
        n = C.c_int(0)
        self.check(self.lib.jdsp_device_count(C.byref(n)))
        return n.value

    def bitrev_table(self, n: int) -> np.ndarray:
        t = np.zeros(n, np.int32)
        self.check(self.lib.jdsp_bitrev_table(C.c_int(n), _ptr(t)))
        return t

    def denoise_params(self, preset: str, mode: int) -> DenoiseParams:
        p = DenoiseParams()
        self.check(self.lib.jdsp_denoise_params_preset(preset.encode(), C.c_int(mode), C.byref(p)))
        return p

    def fastconv_params(self, preset: str) -> FastconvParams:
        p = FastconvParams()
        self.check(self.lib.jdsp_fastconv_params_preset(preset.encode(), C.byref(p)))
        return p

    def mfcc_params(self, preset: str) -> MfccParams:
        p = MfccParams()
        self.check(self.lib.jdsp_mfcc_params_preset(preset.encode(), C.byref(p)))
        return p

    def pitch_params(self, preset: str = "ref") -> PitchParams:
        p = PitchParams()
        self.check(self.lib.jdsp_pitch_params_preset(preset.encode(), C.byref(p)))
        return p

    def mvdr_params(self, preset: str = "ref") -> MvdrParams:
        p = MvdrParams()
        self.check(self.lib.jdsp_mvdr_params_preset(preset.encode(), C.byref(p)))
        return p


class Context:
    """One per host thread per GPU.  `stream` borrows an existing cudaStream_t (e.g. torch's current)."""

    def __init__(self, library: Library | None = None, device: int = 0, stream: int | None = None):
        self.L = library or Library()
        self.lib = self.L.lib
        self.h = C.c_void_p(0)
        if stream is None:
            self.L.check(self.lib.jdsp_create(C.c_int(device), C.byref(self.h)))
        else:
            self.L.check(self.lib.jdsp_create_on_stream(C.c_int(device), C.c_void_p(stream), C.byref(self.h)))

    def close(self) -> None:
        if self.h:
            self.lib.jdsp_destroy(self.h)
            self.h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self) -> None:
        self.L.check(self.lib.jdsp_sync(self.h))

    def cuda_stream(self) -> int:
        return int(self.lib.jdsp_cuda_stream(self.h) or 0)

    # ---- raw device memory and peer mapping (CUDA IPC) for the scatter form of the MFCC kernel ----------------
    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p(0)
        self.L.check(self.lib.jdsp_malloc(self.h, C.byref(p), C.c_size_t(nbytes)))
        return int(p.value)

    def free(self, addr: int) -> None:
        self.L.check(self.lib.jdsp_free(self.h, C.c_void_p(addr)))

    def peer_export(self, addr: int) -> bytes:
        h = (C.c_ubyte * 64)()
        self.L.check(self.lib.jdsp_peer_export(self.h, C.c_void_p(addr), C.byref(h)))
        return bytes(h)

    def peer_open(self, handle: bytes) -> int:
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p(0)
        self.L.check(self.lib.jdsp_peer_open(self.h, C.byref(h), C.byref(p)))
        return int(p.value)

    def peer_close(self, addr: int) -> None:
        self.L.check(self.lib.jdsp_peer_close(self.h, C.c_void_p(addr)))

    def kernel_launches(self) -> int:
        n = C.c_uint64(0)
        self.L.check(self.lib.jdsp_kernel_launches(self.h, C.byref(n)))
        return n.value

    # ---- K1 -------------------------------------------------------------------------------------------
    def fft_process(self, x: np.ndarray, forward: bool) -> np.ndarray:
        """Host drop-in for FFTProcess / fftw_execute: complex128 [..., n] -> complex128, unnormalised."""
        x = np.ascontiguousarray(x, np.complex128)
        out = np.empty_like(x)
        n = x.shape[-1]
        batch = x.size // n
        self.L.check(self.lib.jdsp_fft_process(self.h, _ptr(x), _ptr(out), C.c_int(n), C.c_int(1 if forward else 0),
                                               C.c_long(batch)))
        return out

    def fft_c2c_f32(self, d_in, d_out, n: int, batch: int, forward: bool) -> None:
        self.L.check(self.lib.jdsp_fft_c2c_f32(self.h, _ptr(d_in), _ptr(d_out), C.c_int(n), C.c_long(batch),
                                               C.c_int(1 if forward else 0)))

    def fft_c2c_f32_host(self, h_in, h_out, n: int, batch: int, forward: bool) -> None:
        """complex64 host buffers (numpy or pinned torch), `batch` transforms back to back."""
        self.L.check(self.lib.jdsp_fft_c2c_f32_host(self.h, _ptr(h_in), _ptr(h_out), C.c_int(n), C.c_long(batch),
                                                    C.c_int(1 if forward else 0)))

    def fft_c2c_f64(self, d_in, d_out, n: int, batch: int, forward: bool) -> None:
        self.L.check(self.lib.jdsp_fft_c2c_f64(self.h, _ptr(d_in), _ptr(d_out), C.c_int(n), C.c_long(batch),
                                               C.c_int(1 if forward else 0)))

    # ---- F5 -------------------------------------------------------------------------------------------
    def roundtrip_dev(self, d_in, in_pitch, d_out, out_pitch, d_f32, f32_pitch, n_fft, n_streams, n_blocks) -> None:
        self.L.check(self.lib.jdsp_roundtrip_i16_dev(self.h, _ptr(d_in), C.c_long(in_pitch), _ptr(d_out),
                                                     C.c_long(out_pitch), _ptr(d_f32), C.c_long(f32_pitch),
                                                     C.c_int(n_fft), C.c_long(n_streams), C.c_long(n_blocks)))

    def roundtrip(self, pcm: np.ndarray, n_fft: int) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, np.int16)
        nb = -(-len(pcm) // n_fft)
        out = np.zeros(nb * n_fft, np.int16)
        n_out = C.c_long(0)
        self.L.check(self.lib.jdsp_roundtrip_i16(self.h, _ptr(pcm), C.c_long(len(pcm)), C.c_int(n_fft), _ptr(out),
                                                 C.byref(n_out)))
        return out[: n_out.value]

    def roundtrip_batch_raw(self, h_in, in_pitch, n_streams, n_samples, n_fft, h_out, out_pitch) -> int:
        """Host form on many streams: int16 host rows in, ceil(n/n_fft)*n_fft samples per row out."""
        got = C.c_long(0)
        self.L.check(self.lib.jdsp_roundtrip_batch_i16(self.h, _ptr(h_in), C.c_long(in_pitch), C.c_long(n_streams), C.c_long(n_samples),
                                                       C.c_int(n_fft), _ptr(h_out), C.c_long(out_pitch), C.byref(got)))
        return got.value

    # ---- denoise ----------------------------------------------------------------------------------------
    def denoise_state(self, params: DenoiseParams, n_streams: int) -> "DenoiseState":
        return DenoiseState(self, params, n_streams)

    def denoise(self, x: np.ndarray, params: DenoiseParams) -> np.ndarray:
        """Host form: int16 [n_streams, n_samples] -> int16 [n_streams, (ceil(n/hop)-2)*hop]."""
        x = np.ascontiguousarray(np.atleast_2d(x), np.int16)
        S, n = x.shape
        nb = -(-n // params.hop)
        n_out = max(nb - 2, 0) * params.hop
        out = np.zeros((S, max(n_out, 1)), np.int16)
        got = C.c_long(0)
        self.L.check(self.lib.jdsp_denoise_i16(self.h, C.byref(params), _ptr(x), C.c_long(x.shape[1]), C.c_long(S),
                                               C.c_long(n), _ptr(out), C.c_long(out.shape[1]), C.byref(got)))
        assert got.value == n_out
        return out[:, :n_out]

    def denoise_host_raw(self, params: DenoiseParams, in_ptr, in_pitch, n_streams, n_samples, out_ptr, out_pitch) -> int:
        got = C.c_long(0)
        self.L.check(self.lib.jdsp_denoise_i16(self.h, C.byref(params), _ptr(in_ptr), C.c_long(in_pitch),
                                               C.c_long(n_streams), C.c_long(n_samples), _ptr(out_ptr),
                                               C.c_long(out_pitch), C.byref(got)))
        return got.value

    # ---- pitch (PitchEstimation_method1) ---------------------------------------------------------------------
    def pitch_state(self, params: PitchParams, n_streams: int) -> "PitchState":
        return PitchState(self, params, n_streams)

    def pitch(self, x: np.ndarray, params: PitchParams):
        """Host form: int16 [n_streams, n_samples] -> (arg int32 [n_streams, nb], rmax float64 [n_streams, nb])."""
        x = np.ascontiguousarray(np.atleast_2d(x), np.int16)
        S, n = x.shape
        nb = -(-n // params.block)
        arg = np.zeros((S, max(nb, 1)), np.int32)
        rmax = np.zeros((S, max(nb, 1)), np.float64)
        got = C.c_long(0)
        self.L.check(self.lib.jdsp_pitch_i16(self.h, C.byref(params), _ptr(x), C.c_long(x.shape[1]), C.c_long(S), C.c_long(n),
                                             _ptr(arg), _ptr(rmax), C.byref(got)))
        assert got.value == nb
        return arg[:, :nb], rmax[:, :nb]

    # ---- MVDR beamformer (BeamForming_MVDR_ver1) ---------------------------------------------------------------
    def mvdr_state(self, params: MvdrParams, n_streams: int) -> "MvdrState":
        return MvdrState(self, params, n_streams)

    def mvdr(self, left: np.ndarray, right: np.ndarray, params: MvdrParams) -> np.ndarray:
        """Host form: int16 [n_streams, n_samples] per microphone -> int16 [n_streams, (ceil(n/block)-1)*block]."""
        left = np.ascontiguousarray(np.atleast_2d(left), np.int16)
        right = np.ascontiguousarray(np.atleast_2d(right), np.int16)
        assert left.shape == right.shape
        S, n = left.shape
        nb = -(-n // params.block)
        n_out = max(nb - 1, 0) * params.block
        out = np.zeros((S, max(n_out, 1)), np.int16)
        got = C.c_long(0)
        self.L.check(self.lib.jdsp_mvdr_i16(self.h, C.byref(params), _ptr(left), _ptr(right), C.c_long(left.shape[1]), C.c_long(S),
                                            C.c_long(n), _ptr(out), C.c_long(out.shape[1]), C.byref(got)))
        assert got.value == n_out
        return out[:, :n_out]

    # ---- fast convolution ---------------------------------------------------------------------------------
    def fastconv_state(self, params: FastconvParams, n_sources: int, taps: np.ndarray) -> "FastconvState":
        return FastconvState(self, params, n_sources, taps)

    def fastconv(self, pcm: np.ndarray, taps: np.ndarray, params: FastconvParams) -> np.ndarray:
        """Host form on one source: returns int16 [n_ears, (ceil(n/block)-history)*block]."""
        pcm = np.ascontiguousarray(pcm, np.int16)
        t = np.zeros((params.n_ears, params.n_taps), np.float64)
        taps = np.atleast_2d(np.asarray(taps, np.float64))
        t[:, : min(taps.shape[1], params.n_taps)] = taps[:, : params.n_taps]
        nb = -(-len(pcm) // params.block)
        n_out = max(nb - params.history_blocks, 0) * params.block
        out = np.zeros((params.n_ears, max(n_out, 1)), np.int16)
        got = C.c_long(0)
        self.L.check(self.lib.jdsp_fastconv_i16(self.h, C.byref(params), _ptr(t), _ptr(pcm), C.c_long(len(pcm)), _ptr(out),
                                                C.c_long(out.shape[1]), C.byref(got)))
        assert got.value == n_out
        return out[:, :n_out]

    # ---- MFCC -----------------------------------------------------------------------------------------------
    def mfcc_plan(self, params: MfccParams) -> "MfccPlan":
        return MfccPlan(self, params)

    def mfcc_program(self, pcm: np.ndarray, params: MfccParams) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, np.int16)
        nb = -(-len(pcm) // (2 * params.hop))
        rows = np.zeros((max(2 * nb - 1, 0), params.n_cep), np.float64)
        got = C.c_long(0)
        self.L.check(self.lib.jdsp_mfcc_program_i16(self.h, C.byref(params), _ptr(pcm), C.c_long(len(pcm)), _ptr(rows),
                                                    C.byref(got)))
        assert got.value == rows.shape[0]
        return rows


class PitchState:
    def __init__(self, ctx: Context, params: PitchParams, n_streams: int):
        self.ctx, self.params, self.n_streams = ctx, params, n_streams
        self.h = C.c_void_p(0)
        ctx.L.check(ctx.lib.jdsp_pitch_state_create(ctx.h, C.byref(params), C.c_long(n_streams), C.byref(self.h)))

    def reset(self) -> None:
        self.ctx.L.check(self.ctx.lib.jdsp_pitch_state_reset(self.ctx.h, self.h))

    def close(self) -> None:
        if self.h:
            self.ctx.lib.jdsp_pitch_state_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p(0)

    def run(self, d_in, in_pitch: int, n_blocks: int, d_arg, d_rmax=None) -> None:
        """Device form: d_in int16 [n_streams, in_pitch]; d_arg int32 [n_streams, n_blocks]; d_rmax float64 or None."""
        self.ctx.L.check(self.ctx.lib.jdsp_pitch_i16_dev(self.ctx.h, self.h, _ptr(d_in), C.c_long(in_pitch), C.c_long(n_blocks),
                                                         _ptr(d_arg), _ptr(d_rmax)))


class MvdrState:
    def __init__(self, ctx: Context, params: MvdrParams, n_streams: int):
        self.ctx, self.params, self.n_streams = ctx, params, n_streams
        self.h = C.c_void_p(0)
        ctx.L.check(ctx.lib.jdsp_mvdr_state_create(ctx.h, C.byref(params), C.c_long(n_streams), C.byref(self.h)))

    def reset(self) -> None:
        self.ctx.L.check(self.ctx.lib.jdsp_mvdr_state_reset(self.ctx.h, self.h))

    def close(self) -> None:
        if self.h:
            self.ctx.lib.jdsp_mvdr_state_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p(0)

    def run(self, d_left, d_right, in_pitch: int, n_blocks: int, d_out, out_pitch: int, d_out_f32=None, f32_pitch: int = 0,
            d_vad=None) -> int:
        """Device form; returns the number of blocks emitted per stream."""
        got = C.c_long(0)
        self.ctx.L.check(self.ctx.lib.jdsp_mvdr_i16_dev(self.ctx.h, self.h, _ptr(d_left), _ptr(d_right), C.c_long(in_pitch),
                                                        C.c_long(n_blocks), _ptr(d_out), C.c_long(out_pitch), _ptr(d_out_f32),
                                                        C.c_long(f32_pitch), _ptr(d_vad), C.byref(got)))
        return got.value

    def spatial_corr(self) -> np.ndarray:
        corr = np.zeros((self.n_streams, 2), np.float64)
        self.ctx.L.check(self.ctx.lib.jdsp_mvdr_spatial_corr(self.ctx.h, self.h, _ptr(corr)))
        return corr


class DenoiseState:
    def __init__(self, ctx: Context, params: DenoiseParams, n_streams: int):
        self.ctx, self.params, self.n_streams = ctx, params, n_streams
        self.h = C.c_void_p(0)
        ctx.L.check(ctx.lib.jdsp_denoise_state_create(ctx.h, C.byref(params), C.c_long(n_streams), C.byref(self.h)))

    def reset(self) -> None:
        self.ctx.L.check(self.ctx.lib.jdsp_denoise_state_reset(self.ctx.h, self.h))

    def close(self) -> None:
        if self.h:
            self.ctx.lib.jdsp_denoise_state_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p(0)

    def run(self, d_in, in_pitch: int, n_blocks: int, d_out, out_pitch: int, d_f32=None, f32_pitch: int = 0,
            d_vad=None) -> int:
        got = C.c_long(0)
        self.ctx.L.check(self.ctx.lib.jdsp_denoise_i16_dev(
            self.ctx.h, self.h, _ptr(d_in), C.c_long(in_pitch), C.c_long(n_blocks), _ptr(d_out), C.c_long(out_pitch),
            _ptr(d_f32), C.c_long(f32_pitch), _ptr(d_vad), C.byref(got)))
        return got.value

    def publish_counts(self) -> np.ndarray:
        c = np.zeros(self.n_streams, np.int32)
        self.ctx.L.check(self.ctx.lib.jdsp_denoise_publish_counts(self.ctx.h, self.h, _ptr(c)))
        return c


class FastconvState:
    def __init__(self, ctx: Context, params: FastconvParams, n_sources: int, taps: np.ndarray):
        self.ctx, self.params, self.n_sources = ctx, params, n_sources
        nfilt = 1 if params.shared_filter else n_sources
        t = np.ascontiguousarray(taps, np.float64).reshape(nfilt, params.n_ears, params.n_taps)
        self.h = C.c_void_p(0)
        ctx.L.check(ctx.lib.jdsp_fastconv_state_create(ctx.h, C.byref(params), C.c_long(n_sources), _ptr(t), C.byref(self.h)))

    def reset(self) -> None:
        self.ctx.L.check(self.ctx.lib.jdsp_fastconv_state_reset(self.ctx.h, self.h))

    def close(self) -> None:
        if self.h:
            self.ctx.lib.jdsp_fastconv_state_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p(0)

    def run_host(self, h_in, in_pitch, n_samples, h_out, out_pitch) -> int:
        """Host form: int16 [n_sources][n_samples] in, int16 [n_sources][n_ears][emitted*block] out (ear pitch out_pitch)."""
        got = C.c_long(0)
        self.ctx.L.check(self.ctx.lib.jdsp_fastconv_i16_host(self.ctx.h, self.h, _ptr(h_in), C.c_long(in_pitch), C.c_long(n_samples),
                                                             _ptr(h_out), C.c_long(out_pitch), C.byref(got)))
        return got.value

    def run(self, d_in, in_pitch, n_blocks, d_out, out_pitch, d_f32=None, f32_pitch=0, sources_per_scene: int = 1) -> int:
        got = C.c_long(0)
        if sources_per_scene == 1:
            rc = self.ctx.lib.jdsp_fastconv_i16_dev(self.ctx.h, self.h, _ptr(d_in), C.c_long(in_pitch), C.c_long(n_blocks),
                                                    _ptr(d_out), C.c_long(out_pitch), _ptr(d_f32), C.c_long(f32_pitch),
                                                    C.byref(got))
        else:
            rc = self.ctx.lib.jdsp_fastconv_mix_i16_dev(self.ctx.h, self.h, C.c_int(sources_per_scene), _ptr(d_in),
                                                        C.c_long(in_pitch), C.c_long(n_blocks), _ptr(d_out),
                                                        C.c_long(out_pitch), _ptr(d_f32), C.c_long(f32_pitch), C.byref(got))
        self.ctx.L.check(rc)
        return got.value


class MfccPlan:
    def __init__(self, ctx: Context, params: MfccParams):
        self.ctx, self.params = ctx, params
        self.h = C.c_void_p(0)
        ctx.L.check(ctx.lib.jdsp_mfcc_plan_create(ctx.h, C.byref(params), C.byref(self.h)))

    def close(self) -> None:
        if self.h:
            self.ctx.lib.jdsp_mfcc_plan_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p(0)

    def tables(self):
        nbin = self.params.n_fft // 2
        w = np.zeros(nbin, np.float64)
        ch = np.zeros(nbin, np.int32)
        self.ctx.L.check(self.ctx.lib.jdsp_mfcc_plan_tables(self.h, _ptr(w), _ptr(ch)))
        return w, ch

    def n_frames(self, n_samples: int) -> int:
        p = self.params
        return (n_samples - p.frame_len) // p.hop + 1 if n_samples >= p.frame_len else 0

    def run_host(self, h_in, in_pitch, n_utts, n_samples, h_feat, feat_pitch) -> int:
        got = C.c_long(0)
        self.ctx.L.check(self.ctx.lib.jdsp_mfcc_frames_i16(self.ctx.h, self.h, _ptr(h_in), C.c_long(in_pitch), C.c_long(n_utts),
                                                           C.c_long(n_samples), _ptr(h_feat), C.c_long(feat_pitch), C.byref(got)))
        return got.value

    def run(self, d_in, in_pitch, n_utts, n_samples, d_feat, feat_pitch) -> int:
        got = C.c_long(0)
        self.ctx.L.check(self.ctx.lib.jdsp_mfcc_frames_i16_dev(self.ctx.h, self.h, _ptr(d_in), C.c_long(in_pitch),
                                                               C.c_long(n_utts), C.c_long(n_samples), _ptr(d_feat),
                                                               C.c_long(feat_pitch), C.byref(got)))
        return got.value

    def run_multicast(self, d_in, in_pitch, n_utts, n_samples, mc_dest, feat_pitch) -> int:
        """Scatter form through one NVLink multicast address (where utterance 0 of this call lies in the multicast mapping of the matrix)."""
        got = C.c_long(0)
        self.ctx.L.check(self.ctx.lib.jdsp_mfcc_frames_i16_multicast_dev(self.ctx.h, self.h, _ptr(d_in), C.c_long(in_pitch), C.c_long(n_utts),
                                                                         C.c_long(n_samples), _ptr(mc_dest), C.c_long(feat_pitch), C.byref(got)))
        return got.value

    def run_scatter(self, d_in, in_pitch, n_utts, n_samples, dests, feat_pitch) -> int:
        """Scatter form: every feature row goes to all `dests` (device addresses or tensors: where utterance 0 of this call lies inside
        each destination matrix -- this GPU's and, through Context.peer_open, its peers')."""
        got = C.c_long(0)
        arr = (C.c_void_p * len(dests))(*[_ptr(d) for d in dests])
        self.ctx.L.check(self.ctx.lib.jdsp_mfcc_frames_i16_scatter_dev(self.ctx.h, self.h, _ptr(d_in), C.c_long(in_pitch), C.c_long(n_utts),
                                                                       C.c_long(n_samples), C.c_int(len(dests)), arr, C.c_long(feat_pitch),
                                                                       C.byref(got)))
        return got.value
