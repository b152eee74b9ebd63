"""The band form of the tcgen05 decimator (dec_band_kernel: samples on the M side of the MMA, a band of the taps on the
N side, the outputs a row-block's samples owe to the NEXT row-block accumulated in extra columns and added in the
epilogue).  Forced with set_kernel(4) for every shape it accepts and compared bit for bit with the oracle: streaming
blocks with ragged lengths (history in row -1 of the first tile, ragged last row-block), tile and warp boundaries of
the row hand-over, several channels, full-scale input, 1 / 2 / 3-digit taps, the fused NCO mix."""
import math

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S(built_lib):
    import srcdsp_b200
    return srcdsp_b200


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


SHAPES = [(16, 255, 400), (16, 256, 30000), (8, 63, 2000), (2, 9, 100), (4, 129, 300), (4, 31, 5000), (10, 90, 100000),
          (12, 255, 2 ** 22), (32, 64, 127), (64, 300, 500), (16, 17, 1), (6, 193, 900), (16, 513, 200), (2, 65, 40),
          (16, 1, 3), (2, 2, 50), (8, 9, 1000), (64, 2049, 60)]  # no / one extended output; the longest reach (E = 32) at /64


@pytest.mark.parametrize("M,nt,amp", SHAPES)
def test_band_decimator_sweep(S, corc, M, nt, amp):
    rng = np.random.default_rng(M * 1009 + nt)
    taps = rng.integers(-amp, amp + 1, nt).astype(np.int32)
    taps[nt // 2] = amp
    C = 1 if nt > 2000 else 3  # (keeps the oracle's share of the test short)
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    d.set_kernel(4)
    hs = [None] * C
    G = 32 * M
    al = M * 4 // math.gcd(M, 4)  # device rows must stay 16-byte aligned for the TMA feed: block lengths are multiples of 4
    # blocks: > 2 tiles per channel (63 row-blocks each), exactly one row-block, ragged, a single output row, long again
    for blk, n in enumerate([G * 150 + al * 7, G, G * 63 + al, G * 2 + al * 3, G * 64, G * 200 - al]):
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        got = host(d.step(dev(x)))
        assert d.last_kernel.startswith("dec_band"), d.last_kernel
        for c in range(C):
            e, hs[c] = corc.dec_step(taps, M, x[c], hs[c])
            assert np.array_equal(got[c], e), (M, nt, blk, c, np.argwhere(got[c] != e)[:4])
    for c in range(C):
        assert np.array_equal(d.history(c), hs[c])


@pytest.mark.parametrize("M,nt,n_table", [(16, 255, 4096), (8, 63, 4096), (4, 65, 1024), (2, 33, 256), (32, 255, 8192)])
def test_band_fused_mixer(S, corc, M, nt, n_table):
    rng = np.random.default_rng(M + nt)
    taps = O.design_lowpass_taps(nt, M)
    C = 5
    f = np.array([-0.3217, 0.25, 0.0, -0.9, 0.5005], np.float32)  # odd table steps (period N), a short period, zero
    mix = S.Mixer(n_table, channels=C)
    mix.setFrequency(f)
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    d.set_kernel(4)
    ddc = S.Ddc(mix, d)
    st = [[0, None] for _ in range(C)]
    G = 32 * M
    al = M * 4 // math.gcd(M, 4)
    for blk, n in enumerate([G * 130 + al * 3, G * 5, G * 64 + al * 9, G * 127]):
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        if blk == 2:
            mix.adjustFrequency(0.125, ch=1)  # frequency change in mid-stream (phase continuous)
            f[1] += 0.125
        got = host(ddc.step(dev(x)))
        assert d.last_kernel.startswith("dec_band"), d.last_kernel
        for c in range(C):
            fr = corc.mixer_set_frequency(float(f[c]), n_table)
            m, st[c][0] = corc.mixer_step(x[c], st[c][0], fr, n_table)
            e, st[c][1] = corc.dec_step(taps, M, m, st[c][1])
            assert np.array_equal(got[c], e), (M, nt, blk, c)


def test_band_matches_tma_form_on_many_tiles(S):
    """cfg2 in miniature with every CTA walking many tiles: rings and both accumulators wrap; band form == original form."""
    import torch
    C, M, nt, n = 32, 16, 255, 1 << 21
    taps = O.design_lowpass_taps(nt, M)
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, 0x5EED00BB)
    outs = []
    for kind in (5, 4):
        d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
        d.set_kernel(kind)
        for rep in range(3):
            y = d.step(x)
        torch.cuda.synchronize()
        assert d.last_kernel.startswith("dec_band" if kind == 4 else "dec_tma"), d.last_kernel
        outs.append(y)
    assert torch.equal(outs[0], outs[1])


def test_band_rejects_what_it_cannot_do(S):
    x = np.zeros((4096 * 3, 2), np.int16)
    for M, taps in ((3, [1] * 30), (4, [1] * 200), (16, [1 << 24] * 32)):  # odd ratio; reaches 2 row-blocks back; 4-digit taps
        d = S.FilterDnsamplingFir(M, taps, obsolete=True)
        d.set_kernel(4)
        with pytest.raises(S.SrcDspError) as ei:
            d.step(dev(x[: (len(x) // M) * M]))
        assert ei.value.code == -5


@pytest.mark.parametrize("mix", [False, True])
@pytest.mark.parametrize("out_off", [8, 6])  # 6: output rows not 16-byte aligned -> the scalar store path of the epilogue
def test_band_strided_channel_rows_and_output_views(S, corc, mix, out_off):
    """Channel rows that are views into wider buffers (in_stride > n, out_stride > n_out): the TMA tensor map carries the
    stride, neighbouring memory is not touched, unaligned output rows take the scalar stores."""
    import torch
    rng = np.random.default_rng(31 + out_off)
    C, M, nt, n, pad = 3, 16, 255, 16 * 2016 * 3 + 16 * 40, 64
    taps = O.design_lowpass_taps(nt, M)
    big = torch.from_numpy(rng.integers(-32768, 32768, (C, n + pad, 2)).astype(np.int16)).cuda()
    x = big[:, 32: 32 + n]
    obig = torch.full((C, n // M + 41, 2), 12345, dtype=torch.int16, device="cuda")
    y = obig[:, out_off: out_off + n // M]
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    d.set_kernel(4)
    chain, fs = d, None
    if mix:
        m = S.Mixer(channels=C)
        fs = np.array([-0.25, 0.125, 0.7], np.float32)
        m.setFrequency(fs)
        chain = S.Ddc(m, d)
    chain.step(x, out=y)
    assert d.last_kernel.startswith("dec_band")
    yh, xh, oh = host(y), host(x), host(obig)
    for c in range(C):
        e = xh[c]
        if mix:
            e, _ = corc.mixer_step(e, 0, corc.mixer_set_frequency(float(fs[c])))
        e, _ = corc.dec_step(taps, M, e)
        assert np.array_equal(yh[c], e)
    assert (oh[:, :out_off] == 12345).all() and (oh[:, out_off + n // M:] == 12345).all()


def test_band_time_sliced_stream_equals_sequential(S, corc):
    """Time slices with a warm-up halo on the band form (what bench.py --workload cfg5-style sharding does for /16)."""
    from srcdsp_b200.sharding import time_slices
    rng = np.random.default_rng(77)
    M, nt, n = 16, 255, 16 * 2016 * 8
    taps = O.design_lowpass_taps(nt, M)
    x = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
    exp, _ = corc.dec_step(taps, M, x)
    out = np.zeros_like(exp)
    for sl in time_slices(n, 4, [nt], [M]):
        d = S.FilterDnsamplingFir(M, taps, obsolete=True)
        if sl.warmup:  # 256 samples: less than one row-block, the automatic choice takes the IMAD kernel
            d.step(dev(x[sl.start - sl.warmup: sl.start]))
        d.set_kernel(4)
        out[sl.out_start: sl.out_start + sl.out_length] = host(d.step(dev(x[sl.start: sl.start + sl.length])))
    assert np.array_equal(out, exp)
