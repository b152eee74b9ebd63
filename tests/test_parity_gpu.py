"""GPU parity tests proper: the CUDA path, called through the C ABI (via the ctypes mirror
classes), must be BIT-EXACT against the oracle / the golden fixtures generated from the compiled
reference.  Nothing here reads /root/reference."""
import numpy as np
import pytest

import oracle as O
from cases import load_golden, run_product

pytestmark = pytest.mark.gpu

CASES, GOLD = load_golden()


@pytest.fixture(scope="module")
def S(built_lib):
    import srcdsp_b200
    return srcdsp_b200


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def test_loaded_library_is_in_tree(S):
    """The product library the process loaded is the in-tree CUDA build."""
    from srcdsp_b200 import _capi
    maps = open("/proc/self/maps").read()
    assert _capi.LIB_PATH in maps
    assert S.device_count() >= 1


# ---- golden fixtures: host-pointer flavour, device-pointer flavour, unfused chain ----------------
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_host_buffers(S, case):
    assert np.array_equal(run_product(case), GOLD[case["name"]])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_device_buffers(S, case):
    assert np.array_equal(run_product(case, dev, host), GOLD[case["name"]])


@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "ddc"], ids=lambda c: c["name"])
def test_fused_chain_equals_separate_steps(S, case):
    assert np.array_equal(run_product(case, dev, host, fused=False), GOLD[case["name"]])


# ---- known answers ----------------------------------------------------------------
def const_iq(n, re, im):
    return np.tile(np.array([[re, im]], np.int16), (n, 1))


def test_kats(S):
    d = S.FilterDnsamplingFir(8, [100] * 63, obsolete=True)
    assert tuple(d.step(const_iq(1024, 1000, -500))[-1]) == (1538, -770)
    assert d.coeffScaling == 12
    u = S.FilterUpsamplingFir(8, [4096] * 63 + [0])
    y = u.step(const_iq(64, 1000, -500))
    assert tuple(y[-8]) == (8000, -4000) and tuple(y[-1]) == (7000, -3500)
    assert (u.getLength(), u.getImpLength(), u.getUpsamplingRatio()) == (63, 64, 8)
    u.reset()
    assert tuple(u.step(const_iq(64, 1000, -500), iterator_overload=True)[-8]) == (32767, -32768)
    m = S.Mixer()
    m.setFrequency(0.5)
    assert [tuple(v) for v in m.step(const_iq(5, 16384, 0))] == [(16383, 0), (0, 16383), (-16383, 0), (0, -16383), (16383, 0)]
    m.reset(-0.3217)
    assert m.state()[:2] == (0, 3437)


# ---- random sweeps against the C oracle ------------------------------------------------------------
@pytest.mark.parametrize("M,nt", [(1, 17), (2, 9), (3, 31), (4, 1023), (5, 7), (6, 40), (7, 100), (8, 63),
                                  (10, 90), (12, 255), (16, 255), (16, 256), (32, 64), (64, 300), (8, 1), (1, 1)])
def test_decimator_sweep(S, corc, M, nt):
    rng = np.random.default_rng(M * 10007 + nt)
    taps = rng.integers(-500, 500, nt).astype(np.int32)
    taps[0] = 1000  # sum |c| >= 1
    d = S.FilterDnsamplingFir(M, taps, obsolete=True)
    d.setLeftShiftBy2(1 if nt > 4 else 0)
    ls = 1 if nt > 4 else 0
    h = None
    for blk, nb in enumerate([nt + 37, 2000, 3, 1, 700]):
        n = M * (nb // M + 1)
        x = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
        exp, h = corc.dec_step(taps, M, x, h, ls)
        got = host(d.step(dev(x))) if blk % 2 else d.step(x)
        assert np.array_equal(got, exp), (M, nt, blk)
    assert np.array_equal(d.history(), h)


def test_decimator_int32_taps_wraparound(S, corc):
    """Large int32 taps: the accumulator wraps exactly like the reference's int32_t."""
    rng = np.random.default_rng(3)
    taps = rng.integers(-2**24, 2**24, 40).astype(np.int32)
    x = rng.integers(-32768, 32768, (8 * 300, 2)).astype(np.int16)
    exp, _ = corc.dec_step(taps, 8, x)
    assert np.array_equal(S.FilterDnsamplingFir(8, taps).step(x), exp)


def test_decimator_bank_channels_and_strides(S, corc):
    import torch
    rng = np.random.default_rng(11)
    C, M, nt, n = 5, 8, 63, 8 * 520
    taps = O.design_lowpass_taps(nt, M)
    x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    y1 = d.step(x[:, : n // 2])                      # non-contiguous channel stride on the host
    big = torch.zeros((C, n + 13, 2), dtype=torch.int16, device="cuda")  # odd stride: unaligned path
    big[:, :n] = torch.from_numpy(x).cuda()
    y2 = host(d.step(big[:, n // 2: n]))
    for c in range(C):
        e1, h = corc.dec_step(taps, M, x[c, : n // 2])
        e2, h = corc.dec_step(taps, M, x[c, n // 2:], h)
        assert np.array_equal(y1[c], e1) and np.array_equal(y2[c], e2)
        assert np.array_equal(d.history(c), h)


def test_decimator_reset_state_and_errors(S, corc):
    rng = np.random.default_rng(5)
    taps = O.design_lowpass_taps(63, 8)
    x = rng.integers(-32768, 32768, (800, 2)).astype(np.int16)
    d = S.FilterDnsamplingFir(8, taps, obsolete=True)
    a = d.step(x)
    hist = d.history()
    d.reset()
    assert np.array_equal(d.step(x), a)
    d.set_history(0, hist)                      # checkpoint / migrate a stream
    e, _ = corc.dec_step(taps, 8, x, hist)
    assert np.array_equal(d.step(x), e)
    with pytest.raises(S.SrcDspError) as ei:    # dsptl_dnsampling_filters.h:181
        d.step(x[:801 - 2])
    assert ei.value.code == -2
    with pytest.raises(S.SrcDspError) as ei:    # dsptl_dnsampling_filters.h:122 (new header only)
        S.FilterDnsamplingFir(8, [1] * 63)
    assert ei.value.code == -2
    with pytest.raises(S.SrcDspError) as ei:    # default ctor, no taps yet
        S.FilterDnsamplingFir(8).step(x)
    assert ei.value.code == -5
    # setCoeffs again with the same size keeps the history (history.resize is a no-op)
    d2 = S.FilterDnsamplingFir(8, taps, obsolete=True)
    d2.step(x)
    d2.setCoeffs(taps)
    _, h = corc.dec_step(taps, 8, x)
    e, _ = corc.dec_step(taps, 8, x, h)
    assert np.array_equal(d2.step(x), e)


@pytest.mark.parametrize("L,nt", [(1, 5), (2, 8), (3, 30), (4, 32), (5, 35), (8, 64), (8, 128), (16, 64), (10, 200), (64, 512),
                                  (4, 100), (16, 256), (8, 8), (16, 16)])
def test_upsampler_sweep(S, corc, L, nt):
    rng = np.random.default_rng(L * 977 + nt)
    taps = rng.integers(-6000, 6000, nt).astype(np.int32)
    taps[-2:] = 0
    taps[0] = 77
    for sm in (0, 1):
        u = S.FilterUpsamplingFir(L, taps)
        h = None
        for blk, n in enumerate([1, 300, 5, 1111, 64]):
            x = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
            fl = blk == 4
            exp, h = corc.up_step(taps, L, x, h, fl, sm)
            got = host(u.step(dev(x), flush=fl, iterator_overload=sm == 1)) if blk % 2 else \
                u.step(x, flush=fl, iterator_overload=sm == 1)
            assert np.array_equal(got, exp), (L, nt, sm, blk)
        assert np.array_equal(u.history(), h)


@pytest.mark.parametrize("L,nt", [(8, 64), (8, 128), (4, 32), (16, 64), (16, 400)])
def test_upsampler_blocked_kernel_many_tiles(S, corc, monkeypatch, L, nt):
    """The register-blocked kernel (4 phases x 4 inputs per thread, L in {4, 8, 16}) over several CTAs per channel,
    streaming blocks with ragged lengths and a flush; unaligned output rows take the generic kernel: same result."""
    import torch
    rng = np.random.default_rng(L * 131 + nt)
    taps = O.design_interp_taps(nt, L)
    C = 3
    u = S.FilterUpsamplingFir(L, taps, channels=C)
    g = S.FilterUpsamplingFir(L, taps, channels=C)
    hs = [None] * C
    for blk, n in enumerate([10000 + 3, 4096, 2049, 9000]):
        fl = blk == 3
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        got = host(u.step(dev(x), flush=fl))
        n_out = got.shape[1]
        # the same through an output view at a 4-byte offset (not 16-byte aligned): generic kernel
        big = torch.zeros((C, n_out + 8, 2), dtype=torch.int16, device="cuda")
        g.step(dev(x), flush=fl, out=big[:, 1: 1 + n_out])
        assert np.array_equal(host(big[:, 1: 1 + n_out]), got)
        for c in range(C):
            e, hs[c] = corc.up_step(taps, L, x[c], hs[c], fl, 0)
            assert np.array_equal(got[c], e), (L, nt, blk, c)


@pytest.mark.parametrize("L,nt,amp", [(8, 64, 6000), (8, 96, 30000), (8, 8, 100), (4, 32, 6000), (4, 20, 2 ** 22), (16, 64, 6000),
                                      (16, 224, 127), (32, 480, 6000), (32, 32, 40000), (8, 64, 2 ** 22)])
@pytest.mark.parametrize("sm", [0, 1])
@pytest.mark.parametrize("form", [1, 2])
def test_upsampler_tcgen05_kernel(S, corc, monkeypatch, L, nt, amp, sm, form):
    """The tcgen05 forms of the interpolator (int8 MMAs over 4096 outputs, bias column, low byte plane biased to
    signed; form 1 = taps on the M side, form 2 = taps on the N side, 3 or 4 accumulator slots per output), forced
    for every ratio / length / tap magnitude they accept: streaming blocks with ragged lengths, history, flush, both
    overloads, several channels, full-scale input (saturation of the asymmetric clamp)."""
    monkeypatch.setenv("SRCDSP_UP_TC", "1")
    monkeypatch.setenv("SRCDSP_UP_TC_FORM", str(form))
    rng = np.random.default_rng(L * 7919 + nt + sm)
    taps = rng.integers(-amp, amp + 1, nt).astype(np.int32)
    taps[-2:] = 0
    taps[0] = amp
    C = 3
    u = S.FilterUpsamplingFir(L, taps, channels=C)
    hs = [None] * C
    for blk, n in enumerate([4, 4096 + 76, 8, 1111, 3 * 512 * 4, 64, 5]):
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        fl = blk == 5
        got = host(u.step(dev(x), flush=fl, iterator_overload=sm == 1)) if blk % 2 else u.step(x, flush=fl, iterator_overload=sm == 1)
        if blk % 2:  # device rows: the cp.async feed needs them 16-byte aligned (n % 4 == 0), else a CUDA-core kernel runs
            assert u.last_kernel.startswith(("up_tc2" if form == 2 else "up_tc_") if n % 4 == 0 else "up_fir"), (u.last_kernel, n)
        for c in range(C):
            e, hs[c] = corc.up_step(taps, L, x[c], hs[c], fl, sm)
            assert np.array_equal(got[c], e), (L, nt, sm, form, blk, c)


def test_upsampler_kernel_selection(S, monkeypatch):
    """Automatic choice on a batch that fills the machine: up to 8 taps per phase the register-blocked CUDA-core kernel,
    beyond that the tcgen05 kernel (whose rate does not depend on the filter length)."""
    import torch
    monkeypatch.delenv("SRCDSP_UP_TC", raising=False)
    C, L, n = 128, 8, 1 << 14
    x = torch.zeros((C, n, 2), dtype=torch.int16, device="cuda")
    for nt, want in ((64, "up_fir4"), (96, "up_tc"), (128, "up_fir4")):  # 128 taps: 16 per phase do not fit one MMA
        u = S.FilterUpsamplingFir(L, O.design_interp_taps(nt, L), channels=C)
        u.step(x)
        assert u.last_kernel.startswith(want), (nt, u.last_kernel)


def test_upsampler_tcgen05_many_tiles_match_cuda_core_kernel(S, monkeypatch):
    """cfg-4 shape in miniature with every CTA walking several tiles: the stage ring and both accumulators wrap."""
    import torch
    C, L, nt, n = 64, 8, 64, 1 << 16
    taps = O.design_interp_taps(nt, L)
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, 0x5EED00C4)
    outs = []
    for tc, form in (("0", "1"), ("1", "1"), ("1", "2")):
        monkeypatch.setenv("SRCDSP_UP_TC", tc)
        monkeypatch.setenv("SRCDSP_UP_TC_FORM", form)
        u = S.FilterUpsamplingFir(L, taps, channels=C)
        for rep in range(3):  # carried history + repeated launches
            y = u.step(x)
        torch.cuda.synchronize()
        assert u.last_kernel.startswith(("up_tc2" if form == "2" else "up_tc_") if tc == "1" else "up_fir4")
        outs.append(y)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_upsampler_bank_and_errors(S, corc):
    rng = np.random.default_rng(2)
    C, L, nt = 4, 8, 64
    taps = O.design_interp_taps(nt, L)
    x = rng.integers(-32768, 32768, (C, 777, 2)).astype(np.int16)
    u = S.FilterUpsamplingFir(L, taps, channels=C)
    y = host(u.step(dev(x)))
    for c in range(C):
        e, _ = corc.up_step(taps, L, x[c])
        assert np.array_equal(y[c], e)
    with pytest.raises(S.SrcDspError) as ei:    # upsampling_filters.h:113
        S.FilterUpsamplingFir(8, [1] * 63)
    assert ei.value.code == -2
    with pytest.raises(S.SrcDspError) as ei:    # upsampling_filters.h:155
        S.FilterUpsamplingFir(8).step(x[0])
    assert ei.value.code == -5


@pytest.mark.parametrize("n_table", [4096, 1024, 1000])
def test_mixer_sweep(S, corc, n_table):
    rng = np.random.default_rng(n_table)
    C = 3
    fs = np.array([-0.3217, 0.5, 0.0371], np.float32)
    m = S.Mixer(n_table=n_table, channels=C)
    m.setFrequency(fs)
    phi = [0] * C
    for blk, n in enumerate([1, 4099, 64, 7]):
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        got = host(m.step(dev(x))) if blk % 2 else m.step(x)
        for c in range(C):
            fr = corc.mixer_set_frequency(float(fs[c]), n_table)
            e, phi[c] = corc.mixer_step(x[c], phi[c], fr, n_table)
            assert np.array_equal(got[c], e), (n_table, blk, c)
            assert m.state(c)[:2] == (phi[c], fr)
    xin = dev(rng.integers(-32768, 32768, (C, 512, 2)).astype(np.int16))
    ref_out = host(m.step(xin.clone()))
    for c in range(C):
        m.set_state(c, phi[c], m.state(c)[1], float(fs[c]))
    m.step(xin, out=xin)  # in place
    assert np.array_equal(host(xin), ref_out)
    with pytest.raises(S.SrcDspError):
        m.setFrequency(1.5)  # mixers.h:54


def test_multichannel_ddc_two_stage(S, corc):
    """cfg-3 shape in miniature: per-channel NCO, mix + /8 + /4, streaming blocks, C channels."""
    rng = np.random.default_rng(33)
    C, n = 6, 32 * 150
    t1, t2 = O.design_lowpass_taps(63, 8), O.design_lowpass_taps(63, 4)
    fs = (-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32)
    m = S.Mixer(channels=C)
    m.setFrequency(fs)
    chain = S.Ddc(m, S.FilterDnsamplingFir(8, t1, channels=C, obsolete=True),
                  S.FilterDnsamplingFir(4, t2, channels=C, obsolete=True))
    st = [(0, None, None)] * C
    for blk in range(3):
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        got = host(chain.step(dev(x))) if blk % 2 else chain.step(x)
        for c in range(C):
            phi, h1, h2 = st[c]
            fr = corc.mixer_set_frequency(float(fs[c]))
            y, phi = corc.mixer_step(x[c], phi, fr)
            y, h1 = corc.dec_step(t1, 8, y, h1)
            y, h2 = corc.dec_step(t2, 4, y, h2)
            st[c] = (phi, h1, h2)
            assert np.array_equal(got[c], y), (blk, c)


def test_host_staging_pipeline_many_chunks(S, corc):
    """Host buffers larger than one staging chunk: the chunked H2D/kernel/D2H pipeline must carry
    history across chunks exactly like one call."""
    C, M, nt = 64, 16, 255
    n = 16 * 40000  # 64 ch x 640k samples = 164 MB > the 48 MB staging chunk
    taps = O.design_lowpass_taps(nt, M)
    x = np.stack([corc.synth(0x5EED0002, c, 0, n, 0) for c in range(C)])
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    y = d.step(x)
    for c in (0, 31, 63):
        e, _ = corc.dec_step(taps, M, x[c])
        assert np.array_equal(y[c], e)


# ---- BASELINE-size checks through size-independent properties ----------------------------------------
def _spot_check_dec(y, taps, M, seed, ch_list, n_out, shift, rng, count=400, mixer=None):
    """Recompute randomly chosen outputs from the closed form on the counter-based input."""
    taps64 = np.asarray(taps, np.int64)
    nt = taps64.size
    for _ in range(count):
        c = int(rng.choice(ch_list))
        i = int(rng.integers(0, n_out))
        lo = i * M - (nt - 1)
        xs = O.np_synth(seed, c, max(lo, 0), i * M - max(lo, 0) + 1, 2).astype(np.int64)
        if lo < 0:
            xs = np.concatenate([np.zeros((-lo, 2), np.int64), xs])
        acc = (taps64[::-1, None] * xs).sum(axis=0)
        acc = acc.astype(np.uint64).astype(np.uint32).view(np.int32).astype(np.int64)
        exp = np.clip(acc >> shift, -32767, 32767)
        assert tuple(int(v) for v in y[c, i]) == tuple(int(v) for v in exp), (c, i)


def test_cfg2_full_size_spot_and_split_invariance(S, corc):
    """BASELINE cfg 2 (256 ch x 16 Mi samples, /16, 255 taps) device resident: random outputs are
    recomputed from the closed form, and one call equals two half calls (history carried)."""
    import torch
    C, M, nt, n = 256, 16, 255, 1 << 24
    taps = O.design_lowpass_taps(nt, M)
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, 0x5EED0002, amp_shift=2)
    # device generator == host generator
    assert np.array_equal(host(x[17, 12345:12345 + 64]), corc.synth(0x5EED0002, 17, 12345, 64, 2))
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    y = d.step(x)
    d.reset()
    y2 = torch.empty_like(y)
    d.step(x[:, : n // 2], out=y2[:, : n // 2 // M])
    d.step(x[:, n // 2:], out=y2[:, n // 2 // M:])
    assert torch.equal(y, y2)
    yh = host(y[[0, 100, 255]])
    rng = np.random.default_rng(0)
    sel = {0: yh[0], 100: yh[1], 255: yh[2]}

    class _Y:
        def __getitem__(self, ci):
            return sel[ci[0]][ci[1]]
    _spot_check_dec(_Y(), taps, M, 0x5EED0002, [0, 100, 255], n // M, corc.dec_coeff_scaling(taps), rng, 300)
    # checksum of checksums against the oracle on one full channel prefix
    e, _ = corc.dec_step(taps, M, corc.synth(0x5EED0002, 100, 0, 1 << 18, 2))
    assert np.array_equal(yh[1][: (1 << 18) // M], e)


# ---- tcgen05 int8 Toeplitz kernel (forced with set_kernel(5)) ----------------------------------------
@pytest.mark.parametrize("M,nt,amp", [(16, 255, 400), (16, 256, 30000), (8, 63, 2000), (4, 1023, 300), (2, 9, 100),
                                      (1, 33, 8000), (3, 31, 500), (5, 50, 500), (10, 90, 100000), (12, 255, 2 ** 22),
                                      (32, 64, 127), (64, 300, 500), (16, 17, 1)])
@pytest.mark.parametrize("kind", [5, 3])
def test_tc_decimator_sweep(S, corc, M, nt, amp, kind):
    """The tensor-core path is exact for every ratio / length / tap magnitude it accepts, across
    streaming blocks (history), ragged tile ends and several channels (persistent tile loop)."""
    rng = np.random.default_rng(M * 7919 + nt)
    taps = rng.integers(-amp, amp + 1, nt).astype(np.int32)
    taps[0] = amp
    C = 3
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    d.set_kernel(kind)
    hs = [None] * C
    for blk, n_out in enumerate([4096 * 2 + 77, 5, 4096, 1300, 4096 * 3]):
        n = n_out * M
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        got = host(d.step(dev(x))) if blk % 2 == 0 else d.step(x)
        # kind 2 is fed by TMA whenever a whole row-block (32 outputs) exists and the channel rows are
        # 16-byte aligned, kind 3 never
        assert d.last_kernel.startswith("dec_tma" if kind == 5 and n_out >= 32 and n % 4 == 0 else "dec_tc"), d.last_kernel
        for c in range(C):
            exp, hs[c] = corc.dec_step(taps, M, x[c], hs[c])
            assert np.array_equal(got[c], exp), (M, nt, blk, c)


def test_tc_matches_imad_on_unaligned_buffers(S, corc):
    import torch
    rng = np.random.default_rng(99)
    C, M, nt, n = 4, 16, 255, 16 * 9000
    taps = O.design_lowpass_taps(nt, M)
    big = torch.from_numpy(rng.integers(-32768, 32768, (C, n + 7, 2)).astype(np.int16)).cuda()
    x = big[:, 3: 3 + n]  # 12-byte offset: the 16-byte fast path is off
    outs = []
    for kind in (1, 2, 3):
        d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
        d.set_kernel(kind)
        outs.append(host(d.step(x)))
        assert d.last_kernel.startswith("dec_tc" if kind > 1 else "dec_fir")  # no TMA on unaligned rows
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], outs[2])
    e, _ = corc.dec_step(taps, M, host(x[2]))
    assert np.array_equal(outs[1][2], e)


@pytest.mark.parametrize("w,groups,stages,raw", [(4, 3, 0, 0), (4, 2, 0, 0), (4, 4, 0, 0), (8, 2, 0, 0), (8, 1, 0, 0), (4, 3, 4, 0),
                                                 (4, 2, 5, 4), (4, 2, 3, 3)])
@pytest.mark.parametrize("mix", [False, True])
def test_tma_pipeline_shapes_match_imad(S, monkeypatch, w, groups, stages, raw, mix):
    """Every converter-group / ring-depth shape of the TMA-fed kernel gives the IMAD kernel's output when each
    CTA walks many tiles (7 per CTA here): the rings wrap, groups overtake each other, TMA boxes land out of order."""
    import torch
    C, M, nt, n = 32, 16, 255, 1 << 21
    taps = O.design_lowpass_taps(nt, M)
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, 0x5EED00AA)
    outs = []
    for kind in (1, 5):
        if kind == 5:
            monkeypatch.setenv("SRCDSP_TMA_W", str(w))
            monkeypatch.setenv("SRCDSP_TMA_GROUPS", str(groups))
            if stages:
                monkeypatch.setenv("SRCDSP_TMA_STAGES", str(stages))
            if raw:
                monkeypatch.setenv("SRCDSP_TMA_RAW", str(raw))
        d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
        d.set_kernel(kind)
        chain = d
        if mix:
            m = S.Mixer(channels=C)
            m.setFrequency((-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32))
            chain = S.Ddc(m, d)
        for rep in range(3):  # carried history + repeated launches
            y = chain.step(x)
        torch.cuda.synchronize()
        outs.append(y)
        assert d.last_kernel.startswith("dec_tma" if kind == 5 else "dec_fir")
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("mix", [False, True])
def test_tma_strided_channel_rows_and_output_views(S, corc, mix):
    """Channel rows that are views into wider buffers (in_stride > n, out_stride > n_out, 16-byte aligned):
    the TMA tensor map carries the stride; neighbouring memory is not touched."""
    import torch
    rng = np.random.default_rng(31)
    C, M, nt, n, pad = 3, 16, 255, 16 * 4096 * 3, 64
    taps = O.design_lowpass_taps(nt, M)
    big = torch.from_numpy(rng.integers(-32768, 32768, (C, n + pad, 2)).astype(np.int16)).cuda()
    x = big[:, 32: 32 + n]                               # 128-byte offset, stride n + pad
    obig = torch.full((C, n // M + 40, 2), 12345, dtype=torch.int16, device="cuda")
    y = obig[:, 8: 8 + n // M]
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    d.set_kernel(5)
    chain, fs = d, None
    if mix:
        m = S.Mixer(channels=C)
        fs = np.array([-0.25, 0.125, 0.7], np.float32)
        m.setFrequency(fs)
        chain = S.Ddc(m, d)
    chain.step(x, out=y)
    assert d.last_kernel.startswith("dec_tma")
    yh, xh, oh = host(y), host(x), host(obig)
    for c in range(C):
        e = xh[c]
        if mix:
            e, _ = corc.mixer_step(e, 0, corc.mixer_set_frequency(float(fs[c])))
        e, _ = corc.dec_step(taps, M, e)
        assert np.array_equal(yh[c], e)
    assert (oh[:, :8] == 12345).all() and (oh[:, 8 + n // M:] == 12345).all()


@pytest.mark.parametrize("M,nt,amp", [(4, 1023, 300), (16, 255, 400), (8, 63, 2000), (2, 9, 100), (1, 33, 8000), (3, 31, 500),
                                      (6, 40, 500), (10, 90, 30000), (12, 255, 32767), (32, 64, 127), (16, 17, 1), (2, 700, 32767)])
@pytest.mark.parametrize("mix", [False, True])
def test_tma_p2_geometry(S, corc, monkeypatch, M, nt, amp, mix):
    """The "p2" tile geometry of the TMA-fed kernel (taps of at most 2 signed byte digits: 64 outputs x 2 digit slots per
    row-block, the two byte planes in separate accumulator columns), forced for every ratio / length it accepts -- it is
    chosen automatically only for long filters (4+ lags, cfg 5).  Streaming blocks, ragged ends, several channels."""
    monkeypatch.setenv("SRCDSP_TMA_P2", "1")
    rng = np.random.default_rng(M * 104729 + nt)
    taps = rng.integers(-amp, amp + 1, nt).astype(np.int32)
    taps[0] = amp
    C = 3
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    d.set_kernel(5)
    chain, fs = d, None
    if mix:
        m = S.Mixer(channels=C)
        fs = np.array([-0.3217, 0.5, 0.0371], np.float32)
        m.setFrequency(fs)
        chain = S.Ddc(m, d)
    st = [(0, None)] * C
    # (with the fused mixer only blocks the TMA-fed kernel takes: the register-staged fallback has no room for the
    # sine table next to every master layout of this sweep)
    for blk, n_out in enumerate([4096 * 2 + 76, 4096, 1300, 4096 * 3, 64, 100] if mix else [4096 * 2 + 77, 5, 4096, 1300, 4096 * 3, 64, 100]):
        x = rng.integers(-32768, 32768, (C, n_out * M, 2)).astype(np.int16)
        got = host(chain.step(dev(x))) if blk % 2 == 0 else chain.step(x)
        # a whole row-block is 64 outputs here
        assert d.last_kernel.startswith("dec_tma" if n_out >= 64 and (n_out * M) % 4 == 0 else "dec_tc"), (d.last_kernel, n_out)
        for c in range(C):
            phi, h = st[c]
            y = x[c]
            if mix:
                y, phi = corc.mixer_step(y, phi, corc.mixer_set_frequency(float(fs[c])))
            y, h = corc.dec_step(taps, M, y, h)
            st[c] = (phi, h)
            assert np.array_equal(got[c], y), (M, nt, blk, c)


@pytest.mark.parametrize("M,nt,C,n,mix", [(4, 1023, 8, 1 << 22, False), (4, 1023, 8, 1 << 22, True), (16, 255, 32, 1 << 21, False),
                                          (8, 512, 16, 1 << 22, True)])
@pytest.mark.parametrize("p2", ["0", "1"])
def test_tma_long_filter_many_tiles_match_imad(S, monkeypatch, M, nt, C, n, mix, p2):
    """Long filters (many lags per K-step) with every CTA walking many tiles, in both tile geometries: the rings wrap,
    both accumulator buffers alternate, the result is the IMAD kernel's."""
    import torch
    monkeypatch.setenv("SRCDSP_TMA_P2", p2)
    taps = O.design_lowpass_taps(nt, M)
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, 0x5EED00AB)
    outs = []
    for kind in (1, 5):
        d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
        d.set_kernel(kind)
        chain = d
        if mix:
            m = S.Mixer(channels=C)
            m.setFrequency((-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32))
            chain = S.Ddc(m, d)
        for rep in range(2):  # carried history + repeated launches
            y = chain.step(x)
        torch.cuda.synchronize()
        outs.append(y)
        assert d.last_kernel.startswith("dec_tma" if kind == 5 else "dec_fir")
    assert torch.equal(outs[0], outs[1])


def test_tc_rejects_what_it_cannot_do(S):
    d = S.FilterDnsamplingFir(8, [2 ** 24] * 16, obsolete=True)  # needs 4 signed byte digits
    d.set_kernel(5)
    with pytest.raises(S.SrcDspError) as ei:
        d.step(np.zeros((64, 2), np.int16))
    assert ei.value.code == -5
    d.set_kernel(0)  # automatic selection falls back to the IMAD kernel
    d.step(np.zeros((64, 2), np.int16))


@pytest.mark.parametrize("M,nt,n_table", [(16, 255, 4096), (8, 63, 4096), (4, 200, 1024), (2, 50, 4096), (3, 31, 256)])
@pytest.mark.parametrize("kind", [5, 3])
def test_tc_fused_mixer(S, corc, M, nt, n_table, kind):
    """NCO mix fused into the tensor-core kernel's load stage == Mixer::step then decimator::step."""
    rng = np.random.default_rng(M * 31 + nt)
    taps = O.design_lowpass_taps(nt, M)
    C = 3
    fs = np.array([-0.3217, 0.5, 0.0371], np.float32)
    m = S.Mixer(n_table=n_table, channels=C)
    m.setFrequency(fs)
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    d.set_kernel(kind)
    chain = S.Ddc(m, d)
    st = [(0, None)] * C
    for blk, n_out in enumerate([4096 * 2 + 33, 7, 4096 + 1500, 4096 * 2]):
        x = rng.integers(-32768, 32768, (C, n_out * M, 2)).astype(np.int16)
        if blk == 2:
            m.adjustFrequency(0.0101, ch=1)
            fs[1] = corc.mixer_adjust_nominal(float(fs[1]), 0.0101)
        got = host(chain.step(dev(x))) if blk % 2 == 0 else chain.step(x)
        assert d.last_kernel.startswith("dec_tma" if kind == 5 and n_out >= 32 and (n_out * M) % 4 == 0 else "dec_tc")
        for c in range(C):
            phi, h = st[c]
            y, phi = corc.mixer_step(x[c], phi, corc.mixer_set_frequency(float(fs[c]), n_table), n_table)
            y, h = corc.dec_step(taps, M, y, h)
            st[c] = (phi, h)
            assert np.array_equal(got[c], y), (M, nt, blk, c)
            assert m.state(c)[0] == phi


def test_tc_two_stage_chain(S, corc):
    """cfg-3 shape with the tensor-core kernel in both stages (mix fused into stage 1)."""
    rng = np.random.default_rng(77)
    C, n = 4, 32 * 4096 * 2
    t1, t2 = O.design_lowpass_taps(63, 8), O.design_lowpass_taps(63, 4)
    fs = (-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32)
    m = S.Mixer(channels=C)
    m.setFrequency(fs)
    d1 = S.FilterDnsamplingFir(8, t1, channels=C, obsolete=True)
    d2 = S.FilterDnsamplingFir(4, t2, channels=C, obsolete=True)
    d1.set_kernel(5)
    d2.set_kernel(5)
    chain = S.Ddc(m, d1, d2)
    st = [(0, None, None)] * C
    for blk in range(2):
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        got = host(chain.step(dev(x)))
        for c in range(C):
            phi, h1, h2 = st[c]
            y, phi = corc.mixer_step(x[c], phi, corc.mixer_set_frequency(float(fs[c])))
            y, h1 = corc.dec_step(t1, 8, y, h1)
            y, h2 = corc.dec_step(t2, 4, y, h2)
            st[c] = (phi, h1, h2)
            assert np.array_equal(got[c], y), (blk, c)


def test_filter_fir(S, corc):
    """FilterFir (filters.h) == the M = 1 decimator; setCoeffs clears the history."""
    rng = np.random.default_rng(8)
    taps = O.design_lowpass_taps(33, 4)
    f = S.FilterFir(taps, channels=2)
    h = [None, None]
    for blk, n in enumerate([1000, 37, 5000]):
        x = rng.integers(-32768, 32768, (2, n, 2)).astype(np.int16)
        got = host(f.step(dev(x))) if blk % 2 else f.step(x)
        for c in range(2):
            e, h[c] = corc.fir_step(taps, x[c], h[c])
            assert np.array_equal(got[c], e)
    f.setCoeffs(taps)  # filters.h:96 reset()
    x = rng.integers(-32768, 32768, (2, 500, 2)).astype(np.int16)
    e, _ = corc.fir_step(taps, x[1])
    assert np.array_equal(f.step(x)[1], e)


@pytest.mark.parametrize("kernel", [1, 2, 3])
def test_time_sliced_stream_equals_sequential(S, corc, kernel):
    """cfg-5 shape in miniature: one long stream cut into slices that are processed independently
    (as different GPUs would), each after a warm-up halo with the NCO phase set in closed form,
    gives exactly the sequential result."""
    from srcdsp_b200.sharding import time_slices
    M, nt = 4, 1023
    taps = O.design_lowpass_taps(nt, M)
    n = M * 4096 * 6
    x = corc.synth(0x5EED0005, 0, 0, n, 0)
    f = -0.3217
    fr = corc.mixer_set_frequency(f)
    y, _ = corc.mixer_step(x, 0, fr)
    whole, _ = corc.dec_step(taps, M, y)
    outs = []
    for s in time_slices(n, 3, [nt], [M]):
        m = S.Mixer()
        m.setFrequency(f)
        m.set_state(0, s.nco_phase(0, fr, 4096), fr, f)
        d = S.FilterDnsamplingFir(M, taps, obsolete=True)
        d.set_kernel(kernel)
        chain = S.Ddc(m, d)
        if s.warmup:
            chain.step(dev(x[s.start - s.warmup: s.start]))  # outputs discarded
        outs.append(host(chain.step(dev(x[s.start: s.start + s.length]))))
        assert outs[-1].shape[0] == s.out_length
    assert np.array_equal(np.concatenate(outs), whole)


def test_device_buffers_must_not_overlap(S):
    """In-place filtering through DEVICE pointers is refused (a CTA would read halo samples a neighbour has already
    overwritten); host buffers are staged and may alias, as the reference documents for FilterFir::step."""
    import torch
    taps = O.design_lowpass_taps(33, 4)
    x = torch.zeros((8192, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x.view(1, 8192, 2), 0x5EED0F00)
    f = S.FilterFir(taps)
    with pytest.raises(S.SrcDspError) as ei:
        f.step(x, out=x)
    assert ei.value.code == -1
    d = S.FilterDnsamplingFir(4, taps, obsolete=True)
    with pytest.raises(S.SrcDspError) as ei:
        d.step(x, out=x[:2048])
    assert ei.value.code == -1
    u = S.FilterUpsamplingFir(4, O.design_interp_taps(32, 4))
    with pytest.raises(S.SrcDspError) as ei:
        u.step(x[:1024], out=x[512:512 + 4096])
    assert ei.value.code == -1
    # host buffers, in place (M = 1): same result as out of place
    xh = x.cpu().numpy()
    ref = S.FilterFir(taps).step(xh.copy())
    buf = xh.copy()
    S.FilterFir(taps).step(buf, out=buf)
    assert np.array_equal(buf, ref)
    # the mixer is element-wise: in place is fine on the device
    m = S.Mixer()
    m.setFrequency(0.3)
    y = m.step(x.clone())
    m.reset(0.3)
    z = x.clone()
    m.step(z, out=z)
    assert torch.equal(y, z)


def test_chain_members_survive_the_chain_and_its_leader(S, corc):
    """srcdsp_ddc_create lends dec1's stream to the mixer and dec2; destroying the chain, re-binding dec1's stream
    or destroying dec1 must not leave them with a dangling stream."""
    import torch
    rng = np.random.default_rng(9)
    t1, t2 = O.design_lowpass_taps(63, 8), O.design_lowpass_taps(63, 4)
    x = rng.integers(-32768, 32768, (4096, 2)).astype(np.int16)
    m, d1, d2 = S.Mixer(), S.FilterDnsamplingFir(8, t1, obsolete=True), S.FilterDnsamplingFir(4, t2, obsolete=True)
    m.setFrequency(-0.25)
    fr = corc.mixer_set_frequency(-0.25)
    chain = S.Ddc(m, d1, d2)
    e, phi = corc.mixer_step(x, 0, fr)
    e, h1 = corc.dec_step(t1, 8, e)
    e, h2 = corc.dec_step(t2, 4, e)
    assert np.array_equal(chain.step(x), e)
    with torch.cuda.stream(torch.cuda.Stream()):  # a device-buffer step re-binds dec1 (and with it the members)
        y = chain.step(torch.from_numpy(x).cuda())
        torch.cuda.current_stream().synchronize()
    e1, phi = corc.mixer_step(x, phi, fr)
    e1, h1 = corc.dec_step(t1, 8, e1, h1)
    e1, h2 = corc.dec_step(t2, 4, e1, h2)
    assert np.array_equal(y.cpu().numpy(), e1)
    del chain, d1  # the leader goes first; the mixer and dec2 run on private streams again
    import gc
    gc.collect()
    e2, phi = corc.mixer_step(x, phi, fr)
    assert np.array_equal(m.step(x), e2)
    e3, h2 = corc.dec_step(t2, 4, x, h2)
    assert np.array_equal(d2.step(x), e3)
    d2.setCoeffs(t2)  # cudaStreamSynchronize on its own stream
    m.sync()
