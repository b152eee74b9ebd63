"""Parity at BASELINE.json's FULL sizes, through properties that do not need the oracle to process 16 GiB:

  * split invariance -- one call over the whole batch == two calls over its halves with carried state (history, NCO
    phase): every block boundary, tile boundary and 64-bit offset is exercised on the device against itself;
  * oracle prefix -- the first 2^18 samples of some channels through the C oracle (which is pinned to the compiled
    reference);
  * oracle windows anywhere in the batch -- a window [s0, s1) is regenerated on the host from the counter-based
    generator, the NCO phase at s0 comes from the closed form (phi0 + s0 * freq) mod N (mixers.h:177), the window
    runs through the oracle from zero history and everything behind the chain's delay-line reach must equal the
    device's outputs (dsptl_dnsampling_filters.h:198-205).  Windows sit at the end of the batch (phase wrapped
    thousands of times, largest offsets) and across tile / call boundaries.

cfg2 has its own test in test_parity_gpu.py (test_cfg2_full_size_spot_and_split_invariance)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
N_TABLE = 4096


@pytest.fixture(scope="module")
def S(built_lib):
    import srcdsp_b200
    return srcdsp_b200


def host(t):
    return t.cpu().numpy()


def lo_freqs(C):
    return (-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32)


def oracle_window(corc, seed, ch, s0, s1, stages, f=None, amp_shift=2):
    """Outputs of the chain for input samples [s0, s1) of channel `ch`, computed by the oracle from zero history, and
    the number of leading outputs that depend on samples in front of s0 (to be skipped)."""
    x = corc.synth(seed, ch, s0, s1 - s0, amp_shift)
    if f is not None:
        fr = corc.mixer_set_frequency(float(f))
        x, _ = corc.mixer_step(x, (s0 % N_TABLE) * fr % N_TABLE, fr)
    halo, scale = 0, 1
    for taps, M in stages:
        x, _ = corc.dec_step(taps, M, x)
        halo += (len(taps) - 1) * scale
        scale *= M
    return x, -(-halo // scale)


def check_windows(corc, y, seed, channels, windows, stages, freqs=None, amp_shift=2):
    Mt = int(np.prod([m for _, m in stages]))
    for ch in channels:
        for s0, s1 in windows:
            assert s0 % Mt == 0 and s1 % Mt == 0
            exp, skip = oracle_window(corc, seed, ch, s0, s1, stages, None if freqs is None else freqs[ch], amp_shift)
            if s0 == 0:
                skip = 0  # the batch starts from reset state: the prefix is exact from the first output
            got = host(y[ch, s0 // Mt + skip: s1 // Mt])
            assert np.array_equal(got, exp[skip:]), (ch, s0, s1)


def test_ddc16_full_size(S, corc):
    """north_star's target shape: 256 channels x 16 Mi samples, 256 distinct NCO frequencies, mix fused into /16 x 255 taps."""
    import torch
    C, M, nt, n, seed = 256, 16, 255, 1 << 24, 0x5EED0002
    taps, f = O.design_lowpass_taps(nt, M), lo_freqs(C)
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, seed, amp_shift=2)
    mix = S.Mixer(channels=C)
    mix.setFrequency(f)
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    ddc = S.Ddc(mix, d)
    y = ddc.step(x)
    assert d.last_kernel.startswith("dec_band")  # the band form takes /16
    # phase after 2^24 samples, closed form
    for c in (0, 1, 77, 255):
        fr = corc.mixer_set_frequency(float(f[c]))
        assert mix.state(c)[:2] == ((n % N_TABLE) * fr % N_TABLE, fr)
    # split invariance with carried history AND phase (three ragged parts: tile-unaligned boundaries)
    mix.reset()
    mix.setFrequency(f)
    d.reset()
    y2 = torch.empty_like(y)
    cuts = [0, 5 * 65536 + 16 * 37, n // 2 + 16, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ddc.step(x[:, a:b], out=y2[:, a // M: b // M])
    assert torch.equal(y, y2)
    del y2
    stages = [(taps, M)]
    check_windows(corc, y, seed, [0, 100, 255], [(0, 1 << 17), (n - (1 << 16), n), (n // 2 - 4096, n // 2 + 8192)], stages, f)
    check_windows(corc, y, seed, [3, 200], [(65536 * 9 - 2048, 65536 * 9 + 4096)], stages, f)


def test_cfg3_full_size(S, corc):
    """BASELINE configs[2]: NCO mix + /8 (63 taps) + /4 (63 taps) on 1024 channels x 4 Mi samples."""
    import torch
    C, n, seed = 1024, 1 << 22, 0x5EED0003
    t1, t2, f = O.design_lowpass_taps(63, 8), O.design_lowpass_taps(63, 4), lo_freqs(C)
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, seed, amp_shift=2)
    mix = S.Mixer(channels=C)
    mix.setFrequency(f)
    d1 = S.FilterDnsamplingFir(8, t1, channels=C, obsolete=True)
    d2 = S.FilterDnsamplingFir(4, t2, channels=C, obsolete=True)
    ddc = S.Ddc(mix, d1, d2)
    y = ddc.step(x)
    mix.reset()
    mix.setFrequency(f)
    d1.reset()
    d2.reset()
    y2 = torch.empty_like(y)
    cuts = [0, 32 * 4001, n // 2 + 32, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ddc.step(x[:, a:b], out=y2[:, a // 32: b // 32])
    assert torch.equal(y, y2)
    del y2
    stages = [(t1, 8), (t2, 4)]
    check_windows(corc, y, seed, [0, 511, 1023], [(0, 1 << 16), (n - (1 << 15), n), (n // 2 - 2048, n // 2 + 4096)], stages, f)


def test_cfg4_full_size(S, corc):
    """BASELINE configs[3]: interpolate-by-8, 256 channels, 16 streaming blocks of 2^16 input samples with carried
    history (2 Gi outputs) == one call over the whole 2^20 inputs; oracle prefix and a late window."""
    import torch
    C, L, nt, nb, blocks, seed = 256, 8, 64, 1 << 16, 16, 0x5EED0004
    taps = O.design_interp_taps(nt, L)
    n = nb * blocks
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, seed, amp_shift=2)
    u = S.FilterUpsamplingFir(L, taps, channels=C)
    y = torch.empty((C, n * L, 2), dtype=torch.int16, device="cuda")
    for b in range(blocks):
        u.step(x[:, b * nb:(b + 1) * nb], out=y[:, b * nb * L:(b + 1) * nb * L])
    u.reset()
    y1 = u.step(x)
    assert torch.equal(y, y1)
    del y1
    H = nt // L
    for ch in (0, 128, 255):
        e, _ = corc.up_step(taps, L, corc.synth(seed, ch, 0, 1 << 15, 2))
        assert np.array_equal(host(y[ch, : (1 << 15) * L]), e)
        s0 = n - (1 << 14)
        e, _ = corc.up_step(taps, L, corc.synth(seed, ch, s0, n - s0, 2))  # zero history: skip the first H - 1 inputs' outputs
        assert np.array_equal(host(y[ch, (s0 + H - 1) * L:]), e[(H - 1) * L:])


def test_cfg5_slice_full_size(S, corc):
    """BASELINE configs[4], one slice: decimate-by-4 1023-tap FIR over 2^29 samples of one stream (the 2-digit tile
    geometry, tile offsets beyond 2^31 bytes)."""
    import torch
    M, nt, n, seed = 4, 1023, 1 << 29, 0x5EED0005
    taps = O.design_lowpass_taps(nt, M)
    x = torch.empty((1, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, seed, amp_shift=2)
    d = S.FilterDnsamplingFir(M, taps, channels=1, obsolete=True)
    y = d.step(x)
    assert d.last_kernel.startswith("dec_tma")
    d.reset()
    y2 = torch.empty_like(y)
    cuts = [0, 4 * 1000003, n // 2 + 4, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        d.step(x[:, a:b], out=y2[:, a // M: b // M])
    assert torch.equal(y, y2)
    del y2
    check_windows(corc, y, seed, [0], [(0, 1 << 15), (n - (1 << 15), n), ((1 << 28) - 8192, (1 << 28) + 8192),
                                       (3 * (1 << 27) + 4 * 12345, 3 * (1 << 27) + 4 * 12345 + (1 << 14))], [(taps, M)])
