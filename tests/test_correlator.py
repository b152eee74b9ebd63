"""FixedPatternCorrelator (correlators.h) -- oracle vs the compiled reference (CPU) and the
GPU bank vs the oracle (bit-exact: found flag, corrIndex, bit samples, the 3-deep registers)."""
import os

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "correlator.npz")


def make_case(seed, N, S, n, amp=1500, noise=300, n_bursts=3, full_scale=False):
    """Noise with the pattern embedded (every S-th sample) at a few places; returns (pattern, x, positions)."""
    rng = np.random.default_rng(seed)
    pat = ((rng.integers(0, 2, (N, 2)) * 2 - 1) * amp).astype(np.int32)
    lo, hi = (-32768, 32768) if full_scale else (-noise, noise)
    x = rng.integers(lo, hi, (n, 2)).astype(np.int32)
    pos = sorted(int(p) for p in rng.integers(N * S, n - N * S - 10, n_bursts))
    for p in pos:
        for k in range(N):
            x[p + k * S] += pat[k] * 3 // 2
    return pat, np.clip(x, -32768, 32767).astype(np.int16), pos


def run_blocks(corr, x, blocks):
    """Feeds x in blocks the way a caller of the reference would: after a detection the rest of the block is
    dropped (the reference does not consume it).  Returns the trace of (found, corrIndex, status, bits)."""
    trace, pos = [], 0
    for b in blocks:
        found, idx = corr.step(x[pos: pos + b])
        st = corr.getStatus()
        trace.append((bool(found), int(idx) if found else -1, tuple(st["corrValue"]), tuple(st["energyValue"]),
                      corr.getRefBitSamples().tobytes() if found else b""))
        pos += b
    return trace


CASES = [(1, 32, 4, 6000, [1000, 37, 2963, 2000]), (2, 16, 2, 4000, [4000]), (3, 8, 1, 3000, [5, 1, 2, 992, 2000]),
         (4, 64, 8, 9000, [3000, 3000, 3000])]


@pytest.mark.parametrize("seed,N,S,n,blocks", CASES)
def test_oracle_matches_reference(seed, N, S, n, blocks):
    r = O.ref()
    if r is None:
        pytest.skip("no compiled reference here")
    pat, x, _ = make_case(seed, N, S, n)
    a, b = O.OrcCorrelator(O.corc(), N, S), O.RefCorrelator(r, N, S)
    a.setPattern(pat)
    b.setPattern(pat)
    ta, tb = run_blocks(a, x, blocks), run_blocks(b, x, blocks)
    assert ta == tb
    if N >= 32:
        assert any(t[0] for t in ta)  # the embedded bursts are found (shorter patterns stay below the 2.7 x energy threshold)
    a.reset(), b.reset()
    pat2, x2, _ = make_case(seed + 100, N, S, n, full_scale=True)   # wrap-around arithmetic, no bursts needed
    assert run_blocks(a, x2, blocks) == run_blocks(b, x2, blocks)
    sa, sb = a.getStatus(), b.getStatus()
    assert (sa["coeffsEnergy"], sa["coeffScaling"]) == (sb["coeffsEnergy"], sb["coeffScaling"])


def test_oracle_matches_golden():
    g = np.load(GOLD, allow_pickle=True)
    for key in g.files:
        c = g[key].item()
        a = O.OrcCorrelator(O.corc(), c["N"], c["S"])
        a.setPattern(c["pattern"])
        assert run_blocks(a, c["x"], c["blocks"]) == c["trace"], key


@pytest.fixture(scope="module")
def S_(built_lib):
    import srcdsp_b200
    return srcdsp_b200


# (the blocked scan kernel covers power-of-two strides: tiles of 2048 outputs, tap blocks of 8 with zero padding; other
# strides take the one-output-per-thread kernel)
GPU_CASES = CASES + [(5, 32, 4, 300000, [100000, 100000, 100000]), (6, 20, 16, 9000, [2048, 2049, 4903]),
                     (7, 5, 3, 5000, [5000]), (8, 100, 2, 20000, [7000, 13000]), (9, 33, 1, 10000, [2047, 2, 7951]),
                     (10, 64, 32, 30000, [30000]), (11, 7, 64, 9000, [4500, 4500])]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,N,S,n,blocks", GPU_CASES)
def test_gpu_correlator_matches_oracle(S_, seed, N, S, n, blocks):
    pat, x, _ = make_case(seed, N, S, n, n_bursts=4)
    a = O.OrcCorrelator(O.corc(), N, S)
    g = S_.FixedPatternCorrelator(N, S)
    a.setPattern(pat)
    g.setPattern(pat)
    assert run_blocks(g, x, blocks) == run_blocks(a, x, blocks)
    a.reset(), g.reset()
    pat2, x2, _ = make_case(seed + 100, N, S, n, full_scale=True)
    assert run_blocks(g, x2, blocks) == run_blocks(a, x2, blocks)
    sa, sg = a.getStatus(), g.getStatus()
    assert (sa["coeffsEnergy"], sa["coeffScaling"], sa["thresholdFactor"]) == (sg["coeffsEnergy"], sg["coeffScaling"], sg["thresholdFactor"])


@pytest.mark.gpu
def test_gpu_correlator_bank_and_golden(S_):
    """Several channels at once (device buffers) == one oracle per channel; and the committed reference trace."""
    import torch
    C, N, S, n = 5, 32, 4, 20000
    pats_x = [make_case(50 + c, N, S, n, n_bursts=2) for c in range(C)]
    pat = pats_x[0][0]
    xs = np.stack([px[1] for px in pats_x])
    for c in range(1, C):  # same pattern everywhere: re-embed channel 0's pattern
        for p in pats_x[c][2]:
            for k in range(N):
                xs[c, p + k * S] = np.clip(xs[c, p + k * S].astype(np.int32) + pat[k] * 3 // 2, -32768, 32767)
    g = S_.FixedPatternCorrelator(N, S, channels=C)
    g.setPattern(pat)
    orcs = [O.OrcCorrelator(O.corc(), N, S) for _ in range(C)]
    for o in orcs:
        o.setPattern(pat)
    for lo in range(0, n, 5000):
        f, i = g.step(torch.from_numpy(xs[:, lo: lo + 5000]).cuda())
        for c in range(C):
            ef, ei = orcs[c].step(xs[c, lo: lo + 5000])
            assert bool(f[c]) == ef and (not ef or int(i[c]) == ei)
            assert g.getStatus(c)["corrValue"] == orcs[c].getStatus()["corrValue"]
            if ef:
                assert np.array_equal(g.getRefBitSamples(c), orcs[c].getRefBitSamples())
    gold = np.load(GOLD, allow_pickle=True)
    for key in gold.files:
        c = gold[key].item()
        q = S_.FixedPatternCorrelator(c["N"], c["S"])
        q.setPattern(c["pattern"])
        assert run_blocks(q, c["x"], c["blocks"]) == c["trace"], key


@pytest.mark.gpu
def test_gpu_correlator_errors(S_):
    g = S_.FixedPatternCorrelator(8, 1)
    with pytest.raises(S_.SrcDspError):
        g.step(np.zeros((10, 2), np.int16))                      # no pattern yet
    with pytest.raises(S_.SrcDspError):
        g.setPattern(np.full((8, 2), 20000, np.int32))            # energy assert, correlators.h:183
