"""The single-process multi-device driver (srcdsp_group_*, srcdsp_b200.DdcGroup): channel batches and time slices
with halo must equal the reference's SEQUENTIAL run bit for bit -- outputs, and the state carried into the next
block (dsptl_dnsampling_filters.h:198-205,218-219; mixers.h:177).  On a one-GPU box the members share device 0
(the slicing, warm-up, closed-form phase and state hand-over are the same code); with two or more GPUs the members
sit on different devices, and test_*_on_different_devices refuses to run on fewer."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S(built_lib):
    import srcdsp_b200
    return srcdsp_b200


def n_devices():
    import torch
    return torch.cuda.device_count()


def members(k):
    """k members: on different devices when the box has them, else all on device 0."""
    nd = n_devices()
    return [i % nd for i in range(k)]


def seq_chain(corc, x, state, f, t1, M1, t2=None, M2=0):
    """The reference's sequential chain on one channel; state = [phi, h1, h2]."""
    y = x
    if f is not None:
        y, state[0] = corc.mixer_step(y, state[0], corc.mixer_set_frequency(float(f)))
    y, state[1] = corc.dec_step(t1, M1, y, state[1])
    if M2:
        y, state[2] = corc.dec_step(t2, M2, y, state[2])
    return y


@pytest.mark.parametrize("k", [1, 2, 3, 8])
@pytest.mark.parametrize("shape", ["ddc8x4", "ddc16", "dec4_1023"])
def test_time_slices_equal_the_sequential_run(S, corc, k, shape):
    rng = np.random.default_rng(k * 31 + len(shape))
    if shape == "ddc8x4":
        M1, t1, M2, t2, f = 8, O.design_lowpass_taps(63, 8), 4, O.design_lowpass_taps(63, 4), -0.3217
    elif shape == "ddc16":
        M1, t1, M2, t2, f = 16, O.design_lowpass_taps(255, 16), 0, None, 0.61
    else:
        M1, t1, M2, t2, f = 4, O.design_lowpass_taps(1023, 4), 0, None, None
    Mt = M1 * (M2 or 1)
    g = S.DdcGroup("slices", members(k), 1, M1, t1, M2, t2, n_table=4096 if f is not None else 0)
    if f is not None:
        g.setFrequency(f)
    state = [0, None, None]
    # blocks: long enough to slice, too short to slice (fewer members take part), ragged multiples of Mt
    for blk, n in enumerate([Mt * 4001, Mt * 37, Mt * 1500, Mt, Mt * 2999]):
        x = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
        exp = seq_chain(corc, x, state, f, t1, M1, t2, M2)
        got = g.step(x)
        assert np.array_equal(got, exp), (k, shape, blk)
        lay, used = g.layout()
        assert len(lay) == k and 1 <= used <= k
        if blk == 0 and k > 1:
            assert used == k


@pytest.mark.parametrize("k", [1, 3, 8])
def test_channel_batches_equal_per_channel_objects(S, corc, k):
    rng = np.random.default_rng(k)
    C, M, nt = 10, 16, 255
    taps = O.design_lowpass_taps(nt, M)
    f = (-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32)
    g = S.DdcGroup("channels", members(k), C, M, taps, n_table=4096)
    g.setFrequency(f)
    lay, _ = g.layout()
    assert sum(nc for _, _, nc in lay) == C and [c0 for _, c0, _ in lay] == sorted(c0 for _, c0, _ in lay)
    states = [[0, None, None] for _ in range(C)]
    for blk, n in enumerate([M * 700, M * 3, M * 1111]):
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        got = g.step(x)
        for c in range(C):
            assert np.array_equal(got[c], seq_chain(corc, x[c], states[c], f[c], taps, M)), (k, blk, c)


def test_group_errors(S):
    taps = O.design_lowpass_taps(63, 8)
    g = S.DdcGroup("slices", [0, 0], 1, 8, taps)
    with pytest.raises(S.SrcDspError) as ei:
        g.step(np.zeros((12, 2), np.int16))
    assert ei.value.code == -2
    with pytest.raises(S.SrcDspError) as ei:
        g.setFrequency(0.1)  # no mixer in this group
    assert ei.value.code == -5
    with pytest.raises(S.SrcDspError) as ei:
        S.DdcGroup("channels", [99], 4, 8, taps)
    assert ei.value.code in (-1, -6)


def test_slices_on_different_devices(S, corc):
    """cfg5 in miniature on real hardware: one stream, decimate-by-4 1023-tap FIR, one slice per GPU."""
    nd = n_devices()
    if nd < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch
    rng = np.random.default_rng(5)
    M, taps = 4, O.design_lowpass_taps(1023, 4)
    g = S.DdcGroup("slices", list(range(nd)), 1, M, taps)
    assert sorted(d for d, _, _ in g.layout()[0]) == list(range(nd))
    hin, hout = S.PinnedBuffer(1, 1 << 20), S.PinnedBuffer(1, 1 << 18)
    state = [0, None, None]
    for blk in range(2):
        hin.array[0] = rng.integers(-32768, 32768, (1 << 20, 2)).astype(np.int16)
        g.step(hin.array, out=hout.array)
        assert g.layout()[1] == nd
        assert np.array_equal(hout.array[0], seq_chain(corc, hin.array[0], state, None, taps, M)), blk
    for d in range(nd):
        torch.cuda.synchronize(d)


def test_channel_batches_on_different_devices(S, corc):
    """cfg3 in miniature on real hardware: NCO mix + /8 + /4, channels dealt to the GPUs in contiguous batches."""
    nd = n_devices()
    if nd < 2:
        pytest.skip("needs at least 2 GPUs")
    rng = np.random.default_rng(6)
    C = 4 * nd + 1
    t1, t2 = O.design_lowpass_taps(63, 8), O.design_lowpass_taps(63, 4)
    f = (-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32)
    g = S.DdcGroup("channels", list(range(nd)), C, 8, t1, 4, t2, n_table=4096)
    g.setFrequency(f)
    states = [[0, None, None] for _ in range(C)]
    for blk in range(2):
        x = rng.integers(-32768, 32768, (C, 32 * 2000, 2)).astype(np.int16)
        got = g.step(x)
        for c in range(C):
            assert np.array_equal(got[c], seq_chain(corc, x[c], states[c], f[c], t1, 8, t2, 4)), (blk, c)
