"""Error behaviour and boundary conditions of the C ABI (both backends): invalid arguments return JDSP_ERR_INVALID /
JDSP_ERR_UNSUPPORTED with a message instead of computing garbage; padded row pitches, state reset and multiple
contexts behave."""
import numpy as np
import pytest

from jeicyboodsp_b200 import synth
from jeicyboodsp_b200.binding import SS, Context, JdspError

_CACHE = {}


@pytest.fixture(params=["emul", pytest.param("gpu", marks=pytest.mark.gpu)])
def be(request):
    if request.param not in _CACHE:
        from backends import EmulBackend, GpuBackend
        _CACHE[request.param] = EmulBackend() if request.param == "emul" else GpuBackend()
    return _CACHE[request.param]


def test_invalid_arguments_are_rejected(be):
    z = be.zeros((1, 48), np.complex64)
    with pytest.raises(JdspError) as e:
        be.ctx.fft_c2c_f32(z, z, 48, 1, True)                 # not a power of two
    assert e.value.code == -1
    with pytest.raises(JdspError) as e:
        be.ctx.fft_c2c_f32(be.zeros((1, 1 << 17), np.complex64), be.zeros((1, 1 << 17), np.complex64), 1 << 17, 1, True)
    assert e.value.code == -4                                  # JDSP_ERR_UNSUPPORTED above 2^16
    for bad_n in (1000, 100, 8192):                            # round trip: lengths outside 64..4096 or not powers of two are refused
        with pytest.raises(JdspError) as e:                    # up front (1000 once spun forever building a twiddle table)
            be.ctx.roundtrip(np.zeros(3000, np.int16), bad_n)
        assert e.value.code == -4
    p = be.L.denoise_params("bench", SS)
    p.n_fft = 2048                                             # n_fft != 2*hop
    with pytest.raises(JdspError):
        be.ctx.denoise_state(p, 1)
    p = be.L.denoise_params("bench", SS)
    st = be.ctx.denoise_state(p, 2)
    x = be.zeros((2, 10 * p.hop + 4), np.int16)
    out = be.zeros((2, 8 * p.hop + 4), np.int16)
    with pytest.raises(JdspError) as e:                        # row pitch not a multiple of 8 samples
        st.run(x, 10 * p.hop + 4, 10, out, 8 * p.hop + 4)
    assert "16-byte" in str(e.value)
    st.close()
    with pytest.raises(JdspError):
        be.L.denoise_params("nope", SS)
    with pytest.raises(JdspError):
        be.L.mfcc_params("nope")
    m = be.L.mfcc_params("bench")
    m.hop = 100                                                # not a multiple of 8 samples
    with pytest.raises(JdspError):
        be.ctx.mfcc_plan(m)
    f = be.L.fastconv_params("bench")
    f.n_taps = 400                                             # must be history*block + 1
    with pytest.raises(JdspError):
        be.ctx.fastconv_state(f, 1, np.zeros((1, 2, 400)))


def test_padded_pitches_and_state_reset(be, oracle):
    from oracle.oracle import DenoiseParams as ODP
    p = be.L.denoise_params("bench", SS)
    H, nb, S = p.hop, 40, 2
    x = np.stack([synth.denoise_stream(60 + s, nb * H) for s in range(S)])
    pitch_in, pitch_out = nb * H + 64, (nb - 2) * H + 24      # rows padded beyond the payload
    xin = np.full((S, pitch_in), 12345, np.int16)
    xin[:, : nb * H] = x
    d_in, d_out = be.to_dev(xin), be.zeros((S, pitch_out), np.int16)
    st = be.ctx.denoise_state(p, S)
    for _ in range(2):                                         # second pass after reset must reproduce the first
        st.reset()
        assert st.run(d_in, pitch_in, nb, d_out, pitch_out) == nb - 2
        out = be.to_host(d_out)
        for s in range(S):
            ref = oracle.denoise(x[s], ODP.preset("bench", SS)).out
            assert np.abs(out[s, : (nb - 2) * H].astype(int) - ref.astype(int)).max() <= 1
            assert np.all(out[s, (nb - 2) * H:] == 0)          # padding untouched
    st.close()


def test_two_contexts_are_independent(be):
    ctx2 = Context(be.L, 0)
    rng = np.random.default_rng(3)
    z = rng.normal(size=(2, 512)) + 1j * rng.normal(size=(2, 512))
    a = be.ctx.fft_process(z, True)
    b = ctx2.fft_process(z, True)
    assert np.array_equal(a, b)
    assert ctx2.kernel_launches() >= 1
    ctx2.close()


def test_empty_batches_are_no_ops(be):
    z = be.zeros((1, 64), np.complex64)
    be.ctx.fft_c2c_f32(z, z, 64, 0, True)
    p = be.L.mfcc_params("bench")
    plan = be.ctx.mfcc_plan(p)
    assert plan.n_frames(399) == 0
    feat = be.zeros((1, 1, 13), np.float32)
    assert plan.run(be.zeros((1, 392), np.int16), 392, 1, 392, feat, 13) == 0
    plan.close()


def test_pitch_contract(be, oracle):
    """Pitch entry points: bad presets / framings / lag bounds are rejected, an empty call is a no-op, a padded row pitch
    and a state reset behave, and the host form applies the stale-tail rule to a short final block."""
    with pytest.raises(JdspError):
        be.L.pitch_params("nope")
    p = be.L.pitch_params("ref")
    assert (p.n_fft, p.block, p.min_lag, p.fs) == (1024, 512, 100, 16000.0)
    bad = be.L.pitch_params("ref"); bad.n_fft = 2048
    with pytest.raises(JdspError) as e:
        be.ctx.pitch_state(bad, 1)
    assert e.value.code == -4                                  # JDSP_ERR_UNSUPPORTED framing
    bad = be.L.pitch_params("ref"); bad.min_lag = 511
    with pytest.raises(JdspError):
        be.ctx.pitch_state(bad, 1)
    with pytest.raises(JdspError):
        be.ctx.pitch_state(p, 0)
    H, nb, S, pad = p.block, 6, 3, 24
    x = np.stack([synth.denoise_stream(90 + s, nb * H) for s in range(S)])
    xp = np.zeros((S, nb * H + pad), np.int16); xp[:, : nb * H] = x
    st = be.ctx.pitch_state(p, S)
    arg = be.zeros((S, nb), np.int32)
    st.run(be.to_dev(xp), nb * H + pad, 0, arg, None)          # zero blocks: nothing happens, the keep buffer stays zero
    st.run(be.to_dev(xp), nb * H + pad, nb, arg, None)         # padded rows
    first = be.to_host(arg).copy()
    st.run(be.to_dev(xp), nb * H + pad, nb, arg, None)         # continues: block 0 now pairs with the previous call's last block
    cont = be.to_host(arg).copy()
    st.reset()
    st.run(be.to_dev(xp), nb * H + pad, nb, arg, None)
    again = be.to_host(arg).copy()
    st.close()
    for s in range(S):
        assert np.array_equal(first[s], oracle.pitch(x[s], exact=True)[0])
        assert np.array_equal(cont[s], oracle.pitch(np.concatenate([x[s], x[s]]), exact=True)[0][nb:])
    assert np.array_equal(first, again)
    short = x[0][: 3 * H + 77]                                  # host form, short final block keeps the previous block's tail
    a, r = be.ctx.pitch(short, p)
    ea, er = oracle.pitch(short, exact=True)
    assert a.shape == (1, 4) and np.array_equal(a[0], ea) and np.array_equal(r[0], er)


def test_mvdr_contract(be, oracle):
    """MVDR entry points: bad presets / framings / shapes are rejected, an empty call is a no-op, padded row pitches work and
    reset returns the state to 'no block seen' (the first block emits nothing again)."""
    with pytest.raises(JdspError):
        be.L.mvdr_params("nope")
    p = be.L.mvdr_params("ref")
    assert (p.n_fft, p.block, p.keep, p.energy_thr, p.fs, p.dtime) == (1024, 512, 511, 700.0, 16000.0, 0.0)
    bad = be.L.mvdr_params("ref"); bad.keep = 512
    with pytest.raises(JdspError) as e:
        be.ctx.mvdr_state(bad, 1)
    assert e.value.code == -4
    with pytest.raises(JdspError):
        be.ctx.mvdr_state(p, 0)
    S, nb, B = 2, 6, 512
    pairs = [synth.mvdr_pair(s, nb * B) for s in (1, 2)]
    pitch_in, pitch_out = nb * B + 64, nb * B + 24
    left, right = np.full((S, pitch_in), 4321, np.int16), np.full((S, pitch_in), -77, np.int16)
    for s in range(S):
        left[s, :nb * B], right[s, :nb * B] = pairs[s]
    st = be.ctx.mvdr_state(p, S)
    d_l, d_r, d_out = be.to_dev(left), be.to_dev(right), be.zeros((S, pitch_out), np.int16)
    assert st.run(d_l, d_r, pitch_in, 0, d_out, pitch_out) == 0                   # empty call: nothing happens
    with pytest.raises(JdspError):
        st.run(d_l, d_r, pitch_in - 1, nb, d_out, pitch_out)                      # odd row pitch
    with pytest.raises(JdspError):
        st.run(d_l, d_r, pitch_in, nb, d_out, (nb - 2) * B)                       # output rows too short
    runs = []
    for _ in range(2):
        d_out = be.zeros((S, pitch_out), np.int16)
        assert st.run(d_l, d_r, pitch_in, nb, d_out, pitch_out) == nb - 1
        be.sync()
        runs.append(be.to_host(d_out).copy())
        assert st.run(d_l, d_r, pitch_in, nb, d_out, pitch_out) == nb             # a continued stream emits every block
        be.sync()
        st.reset()
    assert np.array_equal(runs[0], runs[1])
    assert not runs[0][:, (nb - 1) * B:].any()                                    # padding untouched
    for s in range(S):
        d = np.abs(runs[0][s, :(nb - 1) * B].astype(int) - oracle.mvdr(*pairs[s])[0].astype(int))
        assert d.max() <= 1
    st.close()


def test_mvdr_unaligned_rows_take_the_scalar_kernels(be, oracle):
    """Rows that are only 4-byte aligned (input) / 2-byte aligned (output) fall back from the 16-byte-load kernels to the
    scalar-load stats kernel and the two-microphone transform kernel: same decisions, samples within 1 LSB of the oracle."""
    S, nb, B = 3, 10, 512
    pairs = [synth.mvdr_pair(20 + s, nb * B) for s in range(S)]
    pitch_in, pitch_out = nb * B + 2, (nb - 1) * B + 1
    left, right = np.zeros((S, pitch_in), np.int16), np.zeros((S, pitch_in), np.int16)
    for s in range(S):
        left[s, :nb * B], right[s, :nb * B] = pairs[s]
    st = be.ctx.mvdr_state(be.L.mvdr_params("ref"), S)
    d_out, d_vad = be.zeros((S, pitch_out), np.int16), be.zeros((S, nb), np.uint8)
    assert st.run(be.to_dev(left), be.to_dev(right), pitch_in, nb, d_out, pitch_out, None, 0, d_vad) == nb - 1
    be.sync()
    out, vad = be.to_host(d_out), be.to_host(d_vad)
    for s in range(S):
        o_out, _, _, o_vad = oracle.mvdr(*pairs[s])
        assert np.array_equal(vad[s], o_vad)
        assert np.abs(out[s, :(nb - 1) * B].astype(int) - o_out.astype(int)).max() <= 1
    st.close()


@pytest.mark.parametrize("thr", [-1.0, 0.0, 40.0, 700.0, 5000.0, 20000.0])
def test_mvdr_vad_threshold_sweep_clamped_energy_is_exact(be, thr):
    """The 16-byte-load kernels keep the VAD energy in 32 bits by clamping |v| above sqrt(thr N); the scalar-load kernel sums
    in 64 bits.  Both must take identical decisions for any threshold (beyond ~8000 the library itself falls back to 64 bits),
    on quiet, loud and full-scale input."""
    rng = np.random.default_rng(int(abs(thr)) + 5)
    nb, B = 16, 512
    sig = [rng.normal(0, a, nb * B) for a in (3.0, 30.0, 300.0, 3000.0)] + [32767.0 * np.sign(rng.normal(0, 1, nb * B)), np.zeros(nb * B)]
    left = np.stack([np.clip(np.round(x), -32768, 32767).astype(np.int16) for x in sig])
    right = left[::-1].copy()
    S = left.shape[0]
    p = be.L.mvdr_params("ref")
    p.energy_thr = thr
    vads = []
    for pad in (0, 2):                      # pad 2: rows only 4-byte aligned -> scalar-load stats kernel with 64-bit sums
        pitch = nb * B + pad
        l2, r2 = np.zeros((S, pitch), np.int16), np.zeros((S, pitch), np.int16)
        l2[:, :nb * B], r2[:, :nb * B] = left, right
        st = be.ctx.mvdr_state(p, S)
        d_out, d_vad = be.zeros((S, nb * B), np.int16), be.zeros((S, nb), np.uint8)
        assert st.run(be.to_dev(l2), be.to_dev(r2), pitch, nb, d_out, nb * B, None, 0, d_vad) == nb - 1
        be.sync()
        vads.append(be.to_host(d_vad).copy())
        st.close()
    assert np.array_equal(vads[0], vads[1])
    if 0.0 < thr < 5000.0:
        assert vads[0].any() and not vads[0].all()      # the sweep really crosses the threshold
