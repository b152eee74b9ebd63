"""The float instantiation of the decimator, dsptl::FilterDnsamplingFir<complex<float>, complex<float>,
complex<float>, float, M> (dsptl_dnsampling_filters.h:43-220; the only float instantiation of the hot path that the
reference can build).  north_star's tolerance for float samples is 1e-5 relative RMS; the bar here is stricter:
BIT-EXACT (the kernel keeps the reference's tap-order sum of rounded products), asserted with array_equal -- a
relative RMS error of exactly 0.

CPU tests: the C / numpy restatements against the compiled reference and against the golden fixtures generated from
it (tests/golden/make_golden_float.py).  GPU tests: the CUDA path through the C ABI against oracle and fixtures."""
import os
import sys

import numpy as np
import pytest

import oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_float as MGF  # noqa: E402

FCASES = MGF.FCASES
FGOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decf.npz")))
REL_RMS_TOL = 1e-5  # north_star; every comparison below is in fact exact


def rel_rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.sum((a - b) ** 2) / max(np.sum(b ** 2), 1e-300)))


def ftaps(rng, nt, kind):
    if kind == "unity":
        h = np.hamming(nt) * np.sinc((np.arange(nt) - (nt - 1) / 2.0) / 4.0)
        return (h / h.sum()).astype(np.float32)
    if kind == "int":
        return np.round(rng.normal(0, 300, nt)).astype(np.float32)
    return rng.normal(0, 3.0, nt).astype(np.float32)


def run_blocks(step, case):
    x, outs, pos = MGF.finput_for(case), [], 0
    for n in case["blocks"]:
        outs.append(step(x[pos:pos + n]))
        pos += n
    return np.concatenate(outs)


# ---- oracle pinning (CPU) ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,nt", [(8, 63), (16, 255), (4, 1023), (1, 17), (3, 30), (5, 11), (2, 4), (64, 128)])
@pytest.mark.parametrize("kind", ["unity", "int", "frac"])
def test_float_oracle_vs_reference(corc, reflib, M, nt, kind):
    rng = np.random.default_rng(M * 131 + nt)
    t = ftaps(rng, nt, kind)
    ls = 1 if kind == "int" else 0
    ref = O.RefDecF(reflib, M, t, obsolete=True)
    ref.setLeftShiftBy2(ls)
    hc = hn = None
    q = -(-(nt - 1) // M) * M  # the reference needs blocks of at least ntaps - 1 samples (:218-219)
    for blk, n in enumerate([q + M * 40, q + M * 300, q + M * 7]):
        x = rng.uniform(-20000, 20000, (n, 2)).astype(np.float32)
        if blk == 1:
            x = np.round(x)
        e = ref.step(x)
        g, hc = corc.decf_step(t, M, x, hc, ls)
        g2, hn = O.np_decf_step(t, M, x, hn, ls)
        assert np.array_equal(e, g) and np.array_equal(e, g2), (M, nt, kind, blk)
    assert corc.decf_coeff_scaling(t) == O.np_decf_coeff_scaling(t)


def _overflow_case(rng, M, nt):
    """Large-gain taps and unscaled samples: sums beyond +-2^31, where the reference's x86-64 build yields the
    "integer indefinite" 0x80000000 (cvttss2si) for BOTH signs -- positive overflow comes out as -32767."""
    t = rng.uniform(-3e4, 3e4, nt).astype(np.float32)
    x = rng.uniform(-3e4, 3e4, (M * 400, 2)).astype(np.float32)
    x[::7] *= np.float32(40.0)
    return t, x  # finite samples only: see the note on non-finite input in include/srcdsp_b200.h


@pytest.mark.parametrize("M,nt", [(8, 63), (4, 32), (1, 9)])
def test_float_out_of_range_sums_oracle_vs_reference(corc, reflib, M, nt):
    rng = np.random.default_rng(77 + M)
    t, x = _overflow_case(rng, M, nt)
    e = O.RefDecF(reflib, M, t, obsolete=True).step(x)
    g, _ = corc.decf_step(t, M, x)
    g2, _ = O.np_decf_step(t, M, x)
    assert np.array_equal(e, g) and np.array_equal(e, g2)
    indefinite = max(-32767, -((1 << 31) >> (corc.decf_coeff_scaling(t) & 31)))  # 0x80000000 >> shift, clamped
    assert (e == indefinite).sum() > 10  # the case does reach the out-of-range conversion


@pytest.mark.gpu
@pytest.mark.parametrize("M,nt", [(8, 63), (4, 32), (1, 9), (16, 255)])
def test_float_out_of_range_sums_gpu(S, corc, M, nt):
    rng = np.random.default_rng(77 + M)
    t, x = _overflow_case(rng, M, nt)
    exp, _ = corc.decf_step(t, M, x)
    assert np.array_equal(S.FilterDnsamplingFirFloat(M, t, obsolete=True).step(x), exp)


@pytest.mark.parametrize("kind", ["unity", "int", "frac"])
def test_float_fir_oracle_vs_reference(corc, reflib, kind):
    """FilterFir<complex<float>, ...> (filters.h:130-169) in age order is the M = 1 float decimator, for any block
    length (its circular buffer has no minimum block size), including after reset()."""
    rng = np.random.default_rng(11)
    t = ftaps(rng, 33, kind)
    ref, h = O.RefFirF(reflib, t), None
    for n in (1000, 5, 1, 777):
        x = rng.uniform(-20000, 20000, (n, 2)).astype(np.float32)
        e = ref.step(x)
        g, h = corc.decf_step(t, 1, x, h)
        assert np.array_equal(e, g), (kind, n)
    ref.reset()
    x = rng.uniform(-20000, 20000, (100, 2)).astype(np.float32)
    assert np.array_equal(ref.step(x), corc.decf_step(t, 1, x)[0])


def test_float_time_sliced_stream_equals_sequential(corc):
    """one long float stream cut into time slices with a warm-up halo (srcdsp_b200.sharding) gives the
    sequential result bit for bit -- every output's float chain only ever sees its own ntaps-sample window."""
    from srcdsp_b200.sharding import time_slices
    rng = np.random.default_rng(21)
    M, nt, n = 4, 1023, 4 * 3000
    t = ftaps(rng, nt, "unity")
    x = rng.uniform(-20000, 20000, (n, 2)).astype(np.float32)
    whole, _ = corc.decf_step(t, M, x)
    for W in (2, 3, 8):
        outs = []
        for sl in time_slices(n, W, [nt], [M]):
            h = None
            if sl.warmup:
                _, h = corc.decf_step(t, M, x[sl.start - sl.warmup:sl.start], h)
            y, _ = corc.decf_step(t, M, x[sl.start:sl.start + sl.length], h)
            assert y.shape[0] == sl.out_length
            outs.append(y)
        assert np.array_equal(np.concatenate(outs), whole), W


def test_float_coeff_scaling_is_integer_abs(corc):
    """dsptl_dnsampling_filters.h:128-132: abs() on a float tap is ::abs(int) there."""
    assert corc.decf_coeff_scaling(np.full(8, 0.9, np.float32)) == 0x80000000   # sum of int(0.9) = 0: undefined -> INT_MIN
    assert corc.decf_coeff_scaling(np.full(8, 1.9, np.float32)) == 3            # 8 x 1
    assert corc.decf_coeff_scaling(np.full(8, -300.7, np.float32)) == 11        # 8 x 300 = 2400


@pytest.mark.parametrize("case", FCASES, ids=[c["name"] for c in FCASES])
def test_float_oracle_matches_golden(corc, case):
    t, h = MGF.ftaps_for(case), [None]

    def step(x):
        y, h[0] = corc.decf_step(t, case["M"], x, h[0], case["left_shift"])
        return y
    got = run_blocks(step, case)
    assert np.array_equal(got, FGOLD[case["name"]]) and rel_rms(got, FGOLD[case["name"]]) <= REL_RMS_TOL


@pytest.mark.parametrize("case", FCASES, ids=[c["name"] for c in FCASES])
def test_float_golden_is_current_reference_output(reflib, case):
    assert np.array_equal(MGF.run_reference_float(case), FGOLD[case["name"]])


# ---- the CUDA path (GPU) -------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def S(built_lib):
    import srcdsp_b200
    return srcdsp_b200


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", FCASES, ids=[c["name"] for c in FCASES])
@pytest.mark.parametrize("where", ["host", "device"])
def test_float_golden_gpu(S, case, where):
    d = S.FilterDnsamplingFirFloat(case["M"], MGF.ftaps_for(case), obsolete=True)
    d.setLeftShiftBy2(case["left_shift"])
    got = run_blocks((lambda x: d.step(x)) if where == "host" else (lambda x: host(d.step(dev(x)))), case)
    assert np.array_equal(got, FGOLD[case["name"]]), rel_rms(got, FGOLD[case["name"]])
    assert rel_rms(got, FGOLD[case["name"]]) <= REL_RMS_TOL


def quad_applies(M, nt):
    return M in (4, 8, 16, 32) and -(-nt // M) >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("M,nt", [(1, 17), (2, 9), (3, 31), (4, 1023), (5, 7), (6, 40), (7, 100), (8, 63), (10, 90),
                                  (12, 255), (16, 255), (16, 256), (32, 64), (64, 300), (8, 1), (1, 1), (24, 48),
                                  (4, 9), (4, 12), (4, 13), (8, 17), (8, 1000), (16, 33), (16, 2047), (32, 65), (32, 512)])
@pytest.mark.parametrize("kind", ["unity", "int", "frac"])
@pytest.mark.parametrize("shape", ["pairs1", "pairs2", "quad"])
def test_float_decimator_sweep(S, corc, monkeypatch, M, nt, kind, shape):
    """Ragged streaming blocks (also shorter than ntaps - 1), host and device buffers, left shift, every tap family,
    every thread shape of the kernels (1 or 2 output pairs per thread; four outputs per thread where that applies:
    3 rows of taps -- the shortest -- up to taps that no longer fit the parameter space)."""
    if shape == "quad":
        if not quad_applies(M, nt):
            pytest.skip("four outputs per thread need M in {4, 8, 16, 32} and more than 2 M taps")
        monkeypatch.setenv("SRCDSP_DECF_QUAD", "1")  # wherever it applies (the default keeps small ratios on the pair kernel)
        if kind == "frac":
            monkeypatch.setenv("SRCDSP_DECF_CT", "0")  # taps through shared memory instead of the parameter space
    else:
        pairs = 1 if shape == "pairs1" else 2
        monkeypatch.setenv("SRCDSP_DECF_QUAD", "0")
        monkeypatch.setenv("SRCDSP_DECF_PAIRS", str(pairs))
        if pairs == 2 and kind == "frac":
            monkeypatch.setenv("SRCDSP_DECF_BLOCKS", "0")  # the per-chunk code alone (power-of-two ratios default to whole blocks)
    rng = np.random.default_rng(M * 10007 + nt)
    t = ftaps(rng, nt, kind)
    ls = 1 if kind == "int" and nt > 4 else 0
    d = S.FilterDnsamplingFirFloat(M, t, obsolete=True)
    d.setLeftShiftBy2(ls)
    assert d.coeffScaling == corc.decf_coeff_scaling(t)
    assert d.last_kernel.startswith("decf_quad" if shape == "quad" else "decf_fir")
    h = None
    for blk, nb in enumerate([nt + 37, 2600, 3, 1, 700]):
        n = M * (nb // M + 1)
        x = rng.uniform(-30000, 30000, (n, 2)).astype(np.float32)
        exp, h = corc.decf_step(t, M, x, h, ls)
        got = host(d.step(dev(x))) if blk % 2 else d.step(x)
        assert np.array_equal(got, exp), (M, nt, kind, blk, rel_rms(got, exp))


@pytest.mark.gpu
def test_float_filter_fir(S, corc):
    rng = np.random.default_rng(12)
    t = ftaps(rng, 33, "frac")
    f, h = S.FilterFirFloat(t, channels=2), [None, None]
    for blk, n in enumerate((1000, 5, 1, 4096)):
        x = rng.uniform(-20000, 20000, (2, n, 2)).astype(np.float32)
        got = host(f.step(dev(x))) if blk % 2 else f.step(x)
        for c in range(2):
            e, h[c] = corc.decf_step(t, 1, x[c], h[c])
            assert np.array_equal(got[c], e), (blk, c)
    f.setCoeffs(t)  # clears the history (filters.h:96)
    x = rng.uniform(-20000, 20000, (2, 64, 2)).astype(np.float32)
    assert np.array_equal(f.step(x)[0], corc.decf_step(t, 1, x[0])[0])


@pytest.mark.gpu
def test_float_bank_channels_strides_and_errors(S, corc):
    import torch
    rng = np.random.default_rng(5)
    C, M, nt = 5, 8, 64
    t = ftaps(rng, nt, "frac")
    d = S.FilterDnsamplingFirFloat(M, t, channels=C)
    hs = [None] * C
    for n in (M * 300, M * 129):
        x = rng.uniform(-20000, 20000, (C, n, 2)).astype(np.float32)
        big = torch.zeros((C, n + 24, 2), dtype=torch.float32, device="cuda")
        big[:, 3:3 + n] = dev(x)                                  # strided rows, 8-byte (not 16-byte) aligned
        out = torch.zeros((C, n // M + 5, 2), dtype=torch.float32, device="cuda")
        d.step(big[:, 3:3 + n], out=out[:, 1:1 + n // M])
        got = host(out[:, 1:1 + n // M])
        for c in range(C):
            e, hs[c] = corc.decf_step(t, M, x[c], hs[c])
            assert np.array_equal(got[c], e), (n, c)
        assert float(out[:, 0].abs().max()) == 0 and float(out[:, 1 + n // M:].abs().max()) == 0
    d.reset()
    x = rng.uniform(-1000, 1000, (C, M * 64, 2)).astype(np.float32)
    got = d.step(x)
    for c in range(C):
        assert np.array_equal(got[c], corc.decf_step(t, M, x[c])[0])
    with pytest.raises(S.SrcDspError):
        d.step(x[:, :M * 64 - 1])                                  # size not a multiple of M (:181)
    with pytest.raises(S.SrcDspError):
        S.FilterDnsamplingFirFloat(M, t[:nt - 1])                  # taps % M != 0 with the new header (:122)
    with pytest.raises(S.SrcDspError):
        S.FilterDnsamplingFirFloat(M).step(x[0])                   # no coefficients yet
    d.setCoeffs(ftaps(rng, 2 * nt, "unity"))                      # reload on a live object: history resized
    assert d.coeffScaling == 0x80000000


@pytest.mark.gpu
def test_float_many_tiles_split_invariance_and_spot_check(S, corc):
    """A batch that fills the machine (cfg2's filter, 64 channels x 1 Mi samples): one call == two half calls
    (carried history across the block boundary), and random outputs recomputed by the oracle from their windows."""
    import torch
    C, M, nt, n = 64, 16, 255, 1 << 20
    t = MGF.ftaps_for(dict(ntaps=nt, M=M, taps="unity"))
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.rand((C, n, 2), device="cuda", generator=g) - 0.5) * 40000
    a = S.FilterDnsamplingFirFloat(M, t, channels=C, obsolete=True)
    b = S.FilterDnsamplingFirFloat(M, t, channels=C, obsolete=True)
    y = a.step(x)
    y2 = torch.cat([b.step(x[:, :n // 2]), b.step(x[:, n // 2:])], dim=1)
    assert torch.equal(y, y2)
    rng = np.random.default_rng(3)
    for _ in range(24):
        c, i = int(rng.integers(C)), int(rng.integers(nt // M + 1, n // M))
        # oracle with history = the nt - 1 samples in front of x[i * M] and one block of M samples starting there
        hist = host(x[c, i * M - (nt - 1): i * M])
        e, _ = corc.decf_step(t, M, host(x[c, i * M: i * M + M]), hist)
        assert np.array_equal(host(y[c, i]), e[0]), (c, i)


@pytest.mark.gpu
@pytest.mark.parametrize("M,nt", [(4, 63), (8, 64), (16, 255), (32, 97), (16, 1500), (16, 256), (8, 23), (32, 14000)])
@pytest.mark.parametrize("off", [2, 3])
def test_float_quad_kernel_interior_tiles(S, corc, monkeypatch, M, nt, off):
    """Many interior tiles (cp.async staging, padded lane stride) of the four-outputs-per-thread kernel, forced for every
    ratio it supports, on rows that are 16-byte (off = 2) or only 8-byte aligned (off = 3): == the pair kernel on every
    channel, == the oracle on one, with carried history."""
    import torch
    rng = np.random.default_rng(M + nt)
    C, n = 3, M * 6000
    t = ftaps(rng, nt, "frac")
    big = (torch.rand((C, 2 * n + 8, 2), device="cuda") - 0.5) * 30000
    x = big[:, off: off + 2 * n]
    monkeypatch.setenv("SRCDSP_DECF_QUAD", "1")
    q = S.FilterDnsamplingFirFloat(M, t, channels=C, obsolete=True)
    assert q.last_kernel.startswith("decf_quad")
    monkeypatch.setenv("SRCDSP_DECF_QUAD", "0")
    try:
        p = S.FilterDnsamplingFirFloat(M, t, channels=C, obsolete=True)
        assert p.last_kernel.startswith("decf_fir")
    except S.SrcDspError as e:  # the pair kernel keeps two shifted tap copies in shared memory: the longest filters only fit the quad kernel
        assert e.code == -2 and nt > 10000
        p = None
    h = None
    for a, b in ((0, n + M * 6), (n + M * 6, 2 * n)):
        yq = q.step(x[:, a:b])
        if p is not None:
            assert torch.equal(yq, p.step(x[:, a:b]))
        e, h = corc.decf_step(t, M, host(x[1, a:b]), h)
        assert np.array_equal(host(yq[1]), e)


@pytest.mark.gpu
def test_float_longest_filters_fall_back_to_the_kernel_that_fits(S, corc):
    """/4 keeps the pair kernel by default, but its two shifted tap copies stop fitting shared memory before the quad
    kernel's single copy does: 16000 taps run on the quad kernel instead of failing.  Beyond both, E_SIZE."""
    rng = np.random.default_rng(99)
    M, nt = 4, 16000
    t = ftaps(rng, nt, "frac")
    d = S.FilterDnsamplingFirFloat(M, t, obsolete=True)
    assert d.last_kernel.startswith("decf_quad")
    h = None
    for n in (M * 2500, M * 700):
        x = rng.uniform(-30000, 30000, (n, 2)).astype(np.float32)
        e, h = corc.decf_step(t, M, x, h)
        assert np.array_equal(d.step(x), e)
    with pytest.raises(S.SrcDspError) as ei:
        S.FilterDnsamplingFirFloat(M, ftaps(rng, 40000, "frac"), obsolete=True)
    assert ei.value.code == -2
