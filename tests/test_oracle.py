"""CPU tests of the parity oracle: known answers from the reference, the three
restatements against each other, against the compiled reference when present, and against the
committed golden fixtures (generated from the compiled reference)."""
import numpy as np
import pytest

import oracle as O
from cases import load_golden, run_oracle

CASES, GOLD = load_golden()


def const_iq(n, re, im):
    return np.tile(np.array([[re, im]], np.int16), (n, 1))


# ---- known-answer tests produced by the compiled reference during the survey ----
def test_kat_decimator_dc_floor(corc):
    out, _ = corc.dec_step([100] * 63, 8, const_iq(1024, 1000, -500))
    assert tuple(out[-1]) == (1538, -770)  # arithmetic shift floors toward -inf


def test_kat_decimator_impulse(corc):
    taps = [10 * (k + 1) for k in range(63)]
    x = np.zeros((128, 2), np.int16)
    x[0] = (20000, -20000)
    out, _ = corc.dec_step(taps, 8, x)
    exp = [(12, -13), (109, -110), (207, -208), (305, -306), (402, -403), (500, -501), (598, -599), (695, -696), (0, 0)]
    assert [tuple(v) for v in out[:9]] == exp


def test_kat_upsampler_dc(corc):
    taps = [4096] * 63 + [0]
    out, _ = corc.up_step(taps, 8, const_iq(64, 1000, -500))
    last = out[-8:]
    assert all(tuple(v) == (8000, -4000) for v in last[:7]) and tuple(last[7]) == (7000, -3500)
    assert corc.up_length(taps) == 63
    out_it, _ = corc.up_step(taps, 8, const_iq(64, 1000, -500), shift_mode=1)
    assert tuple(out_it[-8]) == (32767, -32768)  # asymmetric limitScale<> clamp
    fl, _ = corc.up_step(taps, 8, const_iq(512, 1000, -500), flush=True)
    assert fl.shape[0] == 4096 + 56


def test_kat_mixer(corc):
    assert corc.mixer_set_frequency(-0.3217) == 3437
    T = corc.mixer_table()
    assert (T[1], T[1024], T[3072]) == (25, 16383, -16383)
    x = np.array([[1, 0], [1000, -500], [-1, 0], [32767, 32767]], np.int16)
    y, _ = corc.mixer_step(x, 0, 0)
    assert [tuple(v) for v in y] == [(0, 0), (999, -500), (-1, 0), (32765, 32765)]
    y, _ = corc.mixer_step(const_iq(5, 16384, 0), 0, corc.mixer_set_frequency(0.5))
    assert [tuple(v) for v in y] == [(16383, 0), (0, 16383), (-16383, 0), (0, -16383), (16383, 0)]
    y, _ = corc.mixer_step(const_iq(3, 16384, 0), 0, corc.mixer_set_frequency(-0.5))
    assert [tuple(v) for v in y] == [(16383, 0), (0, -16383), (-16383, 0)]


def test_mixer_frequency_quantisation(corc):
    for f in [0.0, 1.0, -1.0, 0.5, -0.5, 1e-5, -1e-5, -0.000244, 0.99999, -0.99999, 0.3333, -0.7071]:
        fr = corc.mixer_set_frequency(f)
        assert fr == O.np_mixer_set_frequency(f)
        assert 0 <= fr < 4096


# ---- C restatement vs numpy restatement on random full-scale input --------------------------------
@pytest.mark.parametrize("M,nt", [(8, 63), (16, 255), (4, 1023), (3, 31), (1, 17), (16, 256), (5, 7)])
def test_c_vs_numpy_decimator(corc, M, nt):
    rng = np.random.default_rng(M * 1000 + nt)
    taps = rng.integers(-400, 400, nt).astype(np.int32)
    h = hn = None
    for _ in range(3):
        n = M * int(rng.integers(nt // M + 1, nt // M + 60))
        x = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
        a, h = corc.dec_step(taps, M, x, h, 1)
        b, hn = O.np_dec_step(taps, M, x, hn, 1)
        assert np.array_equal(a, b) and np.array_equal(h, hn)


def test_decimator_short_blocks_equal_one_block(corc):
    """Blocks shorter than ntaps-1 (undefined in the reference) behave like one long call."""
    rng = np.random.default_rng(7)
    taps = rng.integers(-400, 400, 63).astype(np.int32)
    x = rng.integers(-32768, 32768, (8 * 40, 2)).astype(np.int16)
    whole, hw = corc.dec_step(taps, 8, x)
    parts, h = [], None
    for b in range(0, x.shape[0], 16):
        y, h = corc.dec_step(taps, 8, x[b:b + 16], h)
        parts.append(y)
    assert np.array_equal(whole, np.concatenate(parts)) and np.array_equal(hw, h)


@pytest.mark.parametrize("L,nt", [(8, 64), (8, 128), (3, 30), (2, 8), (1, 5), (16, 64)])
def test_c_vs_numpy_upsampler(corc, L, nt):
    rng = np.random.default_rng(L * 1000 + nt)
    taps = rng.integers(-6000, 6000, nt).astype(np.int32)
    taps[-1] = 0
    for sm in (0, 1):
        h = hn = None
        for blk in range(3):
            x = rng.integers(-32768, 32768, (int(rng.integers(1, 200)), 2)).astype(np.int16)
            a, h = corc.up_step(taps, L, x, h, blk == 2, sm)
            b, hn = O.np_up_step(taps, L, x, hn, blk == 2, sm)
            assert np.array_equal(a, b) and np.array_equal(h, hn)


def test_c_vs_numpy_mixer_and_synth(corc):
    x = corc.synth(0x5EED0001, 3, 2**33 + 5, 5000, 0)
    assert np.array_equal(x, O.np_synth(0x5EED0001, 3, 2**33 + 5, 5000, 0))
    assert np.array_equal(corc.synth(1, 0, 0, 100, 2), O.np_synth(1, 0, 0, 100, 2))
    for f in (0.0, -0.3217, 0.5, 0.999):
        fr = corc.mixer_set_frequency(f)
        a, pa = corc.mixer_step(x, 17, fr)
        b, pb = O.np_mixer_step(x, 17, fr)
        assert np.array_equal(a, b) and pa == pb


# ---- restatement vs the compiled reference ------------------------------------------------------
@pytest.mark.parametrize("M,nt", [(8, 63), (16, 255), (4, 1023), (2, 9), (1, 33), (32, 64)])
def test_oracle_vs_reference_decimator(corc, reflib, M, nt):
    rng = np.random.default_rng(nt)
    taps = rng.integers(-300, 300, nt).astype(np.int32)
    d = O.RefDecimator(reflib, M, taps)
    d.setLeftShiftBy2(1)
    h = None
    for _ in range(3):
        n = M * ((nt + int(rng.integers(0, 700))) // M + 1)
        x = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
        y, h = corc.dec_step(taps, M, x, h, 1)
        assert np.array_equal(d.step(x), y)


def test_reference_new_header_asserts_on_taps(reflib):
    with pytest.raises(ValueError):
        O.RefDecimator(reflib, 8, [1] * 63, variant=1)  # dsptl_dnsampling_filters.h:122
    a = O.RefDecimator(reflib, 8, [100] * 64, variant=1).step(const_iq(1024, 1000, -500))
    b = O.RefDecimator(reflib, 8, [100] * 64, variant=0).step(const_iq(1024, 1000, -500))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("L,nt", [(8, 64), (8, 128), (3, 30), (16, 64)])
def test_oracle_vs_reference_upsampler(corc, reflib, L, nt):
    rng = np.random.default_rng(nt + L)
    taps = rng.integers(-6000, 6000, nt).astype(np.int32)
    taps[-2:] = 0
    for sm in (0, 1):
        u = O.RefUpsampler(reflib, L, taps)
        assert u.getLength() == corc.up_length(taps) and u.getImpLength() == nt
        h = None
        for blk in range(3):
            x = rng.integers(-32768, 32768, (int(rng.integers(1, 300)), 2)).astype(np.int16)
            y, h = corc.up_step(taps, L, x, h, blk == 2, sm)
            assert np.array_equal(u.step(x, blk == 2, sm), y)


def test_oracle_vs_reference_mixer_fir(corc, reflib):
    rng = np.random.default_rng(5)
    x = rng.integers(-32768, 32768, (3000, 2)).astype(np.int16)
    for f in (0.0, 0.5, -0.5, -0.3217, 1.0, -1.0, -1e-5):
        m = O.RefMixer(reflib)
        m.setFrequency(f)
        fr = corc.mixer_set_frequency(f)
        assert fr == m.freq
        y, phi = corc.mixer_step(x, 0, fr)
        assert np.array_equal(m.step(x), y) and phi == m.phi
    taps = rng.integers(-300, 300, 33).astype(np.int32)
    fir, h = O.RefFir(reflib, taps), None
    for _ in range(2):
        y, h = corc.fir_step(taps, x[:500], h)
        assert np.array_equal(fir.step(x[:500]), y)


# ---- golden fixtures (generated from the compiled reference, tests/golden/make_golden.py) -------
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_golden(corc, case):
    assert np.array_equal(run_oracle(case), GOLD[case["name"]])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_is_current_reference_output(reflib, case):
    import make_golden as MG
    assert np.array_equal(MG.run_reference(case), GOLD[case["name"]])


# ---- live objects: setCoeffs / setCoefficients after samples have been filtered --------------------------------
def _live_script(rng, sizes, unit, n_ops=14):
    """(kind, payload) operations: steps of random length (multiples of `unit`) interleaved with coefficient
    changes that grow, shrink and keep the tap count."""
    ops = []
    for i in range(n_ops):
        if i % 3 == 2:
            ops.append(("coeffs", rng.integers(-3000, 3000, int(rng.choice(sizes))).astype(np.int32)))
        else:
            ops.append(("step", rng.integers(-32768, 32768, (unit * int(rng.integers(1, 40)), 2)).astype(np.int16)))
    return ops


@pytest.mark.parametrize("M,sizes", [(4, (8, 16, 32, 64)), (8, (8, 64, 24)), (1, (5, 33, 2)), (16, (256, 16, 64))])
def test_live_setcoeffs_decimator_vs_reference(reflib, M, sizes):
    """history.resize() on a live object (dsptl_dnsampling_filters.h:127): the numpy object model used by the GPU
    tests equals the compiled reference, whatever the order of growing / shrinking / same-size changes."""
    rng = np.random.default_rng(100 + M)
    ops = _live_script(rng, sizes, M * 8)  # blocks of at least ntaps - 1 are not needed: same-size rule only for the reference
    t0 = rng.integers(-3000, 3000, sizes[0]).astype(np.int32)
    r, m = O.RefDecimator(reflib, M, t0, variant=1), O.NpDecimatorLive(M, t0)
    for i, (kind, v) in enumerate(ops):
        if kind == "coeffs":
            r.setCoeffs(v)
            m.setCoeffs(v)
            if i % 2:
                r.setLeftShiftBy2(1)
                m.setLeftShiftBy2(1)
        else:
            # the reference reads out of bounds for blocks shorter than ntaps - 1 (:218-219): feed at least that
            need = m.taps.size - 1
            if v.shape[0] < need:
                v = np.concatenate([v] * (need // v.shape[0] + 1))
                v = v[: (v.shape[0] // M) * M]
            assert np.array_equal(r.step(v), m.step(v)), (M, i)


@pytest.mark.parametrize("L,sizes", [(8, (64, 128, 8, 32)), (4, (32, 4, 64)), (2, (16, 6, 2)), (16, (64, 256, 16))])
def test_live_setcoefficients_upsampler_vs_reference(reflib, L, sizes):
    """buffer.resize() with `top` left alone (upsampling_filters.h:118): model == compiled reference.  A change that
    would leave `top` outside the new buffer (the reference then writes out of bounds) is skipped by reset()."""
    rng = np.random.default_rng(200 + L)
    ops = _live_script(rng, sizes, 1, n_ops=20)
    t0 = rng.integers(-3000, 3000, sizes[0]).astype(np.int32)
    t0[-1] = 0
    r, m = O.RefUpsampler(reflib, L, t0), O.NpUpsamplerLive(L, t0)
    for i, (kind, v) in enumerate(ops):
        if kind == "coeffs":
            if m.top >= v.size // L:
                r.reset()
                m.reset()
            r.setCoefficients(v)
            m.setCoefficients(v)
            assert (r.getLength(), r.getImpLength()) == (m.getLength(), m.getImpLength())
        else:
            fl = i % 4 == 1
            assert np.array_equal(r.step(v, flush=fl, shift_mode=i % 2), m.step(v, flush=fl, shift_mode=i % 2)), (L, i)


# ---- randomised: C oracle == numpy restatement == compiled reference over random shapes ------------------------------
RATIOS = (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20, 24, 32, 64)  # the instantiations oracle/ref_harness.cpp holds


def _rand_int_taps(rng, nt):
    """Random taps whose full-scale worst case stays inside int32 (sum |c| <= 65535: signed overflow is undefined in the
    reference), with 1-, 2- and 3-byte magnitudes and sparse patterns."""
    amp = int(rng.choice([3, 100, 4000, 60000]))
    t = rng.integers(-amp, amp + 1, nt).astype(np.int64)
    if rng.integers(3) == 0:
        t[rng.random(nt) < 0.6] = 0
    if not t.any():
        t[int(rng.integers(nt))] = amp
    s = int(np.abs(t).sum())
    if s > 65535:
        t = (t * 65535) // s
        if not t.any():
            t[nt // 2] = 1
    return t.astype(np.int32)


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_oracle_vs_reference(corc, reflib, seed):
    """40 random scripts per seed: decimator (both headers where the tap count allows, left shift, 1-3 ragged streaming
    blocks with carried history), upsampler (both overloads, flush, trailing zero taps), mixer (random frequencies with
    adjustFrequency in mid-stream), float decimator -- C oracle == numpy restatement == the unmodified reference."""
    rng = np.random.default_rng(1000 + seed)
    for case in range(40):
        kind = int(rng.integers(4))
        if kind == 0:
            M = int(rng.choice(RATIOS))
            nt = int(rng.integers(1, 400))
            if rng.integers(3) == 0:
                nt = M * max(1, nt // M)  # a multiple of M: the new header accepts it
            taps = _rand_int_taps(rng, nt)
            ls = int(rng.integers(0, 3)) if corc.dec_coeff_scaling(taps) >= 2 else 0
            variant = 1 if nt % M == 0 and rng.integers(2) else 0
            d = O.RefDecimator(reflib, M, taps, variant=variant)
            d.setLeftShiftBy2(ls)
            h = hn = None
            for _ in range(int(rng.integers(1, 4))):
                n = M * ((nt + int(rng.integers(0, 900))) // M + 1)  # >= ntaps - 1 (shorter blocks read out of bounds in the reference)
                x = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
                y, h = corc.dec_step(taps, M, x, h, ls)
                yn, hn = O.np_dec_step(taps, M, x, hn, ls)
                assert np.array_equal(d.step(x), y) and np.array_equal(yn, y), ("dec", seed, case, M, nt, ls, variant)
        elif kind == 1:
            L = int(rng.choice(RATIOS))
            H = int(rng.integers(1, 17))
            taps = _rand_int_taps(rng, L * H)
            if rng.integers(2):
                taps[-int(rng.integers(1, L * H)):] = 0 if L * H > 1 else taps[-1:]
                if not taps.any():
                    taps[0] = 1
            u = O.RefUpsampler(reflib, L, taps)
            assert u.getLength() == corc.up_length(taps)
            h = None
            nblk = int(rng.integers(1, 4))
            for blk in range(nblk):
                x = rng.integers(-32768, 32768, (int(rng.integers(1, 400)), 2)).astype(np.int16)
                flush, sm = (blk == nblk - 1 and bool(rng.integers(2))), int(rng.integers(2))
                y, h = corc.up_step(taps, L, x, h, flush, sm)
                assert np.array_equal(u.step(x, flush, sm), y), ("up", seed, case, L, H, blk, flush, sm)
        elif kind == 2:
            n_table = int(rng.choice([256, 1024, 4096, 8192]))
            m = O.RefMixer(reflib, n_table)
            f = float(np.float32(rng.uniform(-1, 1)))
            m.setFrequency(f)
            fr, phi = corc.mixer_set_frequency(f, n_table), 0
            assert fr == m.freq
            for blk in range(3):
                x = rng.integers(-32768, 32768, (int(rng.integers(1, 3000)), 2)).astype(np.int16)
                y, phi = corc.mixer_step(x, phi, fr, n_table)
                assert np.array_equal(m.step(x), y) and phi == m.phi, ("mixer", seed, case, f, n_table, blk)
        else:
            M = int(rng.choice(RATIOS))
            nt = int(rng.integers(1, 300))
            tk = int(rng.integers(3))
            t = (rng.normal(0, 1, nt) / nt if tk == 0 else rng.integers(-200, 201, nt) if tk == 1 else rng.uniform(-3, 3, nt)).astype(np.float32)
            d = O.RefDecF(reflib, M, t, obsolete=True)
            h = None
            for _ in range(int(rng.integers(1, 3))):
                n = M * ((nt + int(rng.integers(0, 600))) // M + 1)
                x = rng.uniform(-30000, 30000, (n, 2)).astype(np.float32)
                y, h = corc.decf_step(t, M, x, h)
                assert np.array_equal(d.step(x), y), ("decf", seed, case, M, nt, tk)
