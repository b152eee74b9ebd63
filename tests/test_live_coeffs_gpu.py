"""GPU parity of the state surgery the reference does when coefficients are loaded into a LIVE object --
`history.resize()` in FilterDnsamplingFir::setCoeffs (dsptl_dnsampling_filters.h:114-134) and `buffer.resize()` with
`top` left alone in FilterUpsamplingFir::setCoefficients (upsampling_filters.h:107-126) -- and of the golden inputs
against the compiled reference ITSELF when oracle/_ref travelled to the GPU box (it does: built in the container,
git-ignored, not gpurun-ignored).  The object models (oracle.NpDecimatorLive / NpUpsamplerLive) are pinned against
the compiled reference by tests/test_oracle.py::test_live_*."""
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


def _script(rng, sizes, unit, n_ops, C):
    ops = []
    for i in range(n_ops):
        if i % 3 == 2:
            ops.append(("coeffs", rng.integers(-3000, 3000, int(rng.choice(sizes))).astype(np.int32)))
        else:
            ops.append(("step", rng.integers(-32768, 32768, (C, unit * int(rng.integers(1, 40)), 2)).astype(np.int16)))
    return ops


@pytest.mark.parametrize("M,sizes", [(4, (8, 16, 32, 64)), (8, (8, 64, 24)), (1, (5, 33, 2)), (16, (256, 16, 64))])
@pytest.mark.parametrize("device_buffers", [False, True])
def test_live_setcoeffs_decimator(S, M, sizes, device_buffers):
    """setCoeffs() with a different tap count on a decimator that has history: the carried samples stay at the
    FRONT of the resized vector (so they age by the growth), new entries are zero, leftShift is cleared."""
    import torch
    rng = np.random.default_rng(300 + M)
    C = 3
    ops = _script(rng, sizes, M * 8, 14, C)
    t0 = rng.integers(-3000, 3000, sizes[0]).astype(np.int32)
    d = S.FilterDnsamplingFir(M, t0, channels=C)
    models = [O.NpDecimatorLive(M, t0) for _ in range(C)]
    ref = O.ref()
    refs = [O.RefDecimator(ref, M, t0, variant=1) for _ in range(C)] if ref is not None else None
    for i, (kind, v) in enumerate(ops):
        if kind == "coeffs":
            d.setCoeffs(v)
            for m in models:
                m.setCoeffs(v)
            for r in refs or []:
                r.setCoeffs(v)
            if i % 2:
                d.setLeftShiftBy2(1)
                for m in models + (refs or []):
                    m.setLeftShiftBy2(1)
            continue
        need = models[0].taps.size - 1  # the reference needs blocks of at least ntaps - 1 samples (:218-219)
        if v.shape[1] < need:
            v = np.concatenate([v] * (need // v.shape[1] + 1), axis=1)
            v = np.ascontiguousarray(v[:, : (v.shape[1] // M) * M])
        got = d.step(torch.from_numpy(v).cuda()).cpu().numpy() if device_buffers else d.step(v)
        for c in range(C):
            assert np.array_equal(got[c], models[c].step(v[c])), (M, i, c)
            if refs:
                assert np.array_equal(got[c], refs[c].step(v[c])), (M, i, c, "compiled reference")
    for c in range(C):
        assert np.array_equal(d.history(c), models[c].history)


@pytest.mark.parametrize("L,sizes", [(8, (64, 128, 8, 32)), (4, (32, 4, 64)), (2, (16, 6, 2)), (16, (64, 256, 16))])
def test_live_setcoefficients_upsampler(S, L, sizes):
    """setCoefficients() on an interpolator in mid-stream: the raw circular buffer is resized around an unchanged
    `top`; a change that would leave `top` outside the new buffer (an out-of-bounds write in the reference) is
    refused with SRCDSP_E_STATE and leaves the object as it was."""
    rng = np.random.default_rng(400 + L)
    C = 2
    ops = _script(rng, sizes, 1, 24, C)
    t0 = rng.integers(-3000, 3000, sizes[0]).astype(np.int32)
    t0[-1] = 0
    u = S.FilterUpsamplingFir(L, t0, channels=C)
    models = [O.NpUpsamplerLive(L, t0) for _ in range(C)]
    ref = O.ref()
    refs = [O.RefUpsampler(ref, L, t0) for _ in range(C)] if ref is not None else None
    refused = 0
    for i, (kind, v) in enumerate(ops):
        if kind == "coeffs":
            if models[0].top >= v.size // L:
                with pytest.raises(S.SrcDspError) as ei:
                    u.setCoefficients(v)
                assert ei.value.code == -5
                refused += 1
                u.reset()
                for m in models + (refs or []):
                    m.reset()
            u.setCoefficients(v)
            for m in models + (refs or []):
                m.setCoefficients(v)
            assert (u.getLength(), u.getImpLength()) == (models[0].getLength(), models[0].getImpLength())
            continue
        fl, sm = i % 4 == 1, i % 2
        got = u.step(v, flush=fl, iterator_overload=sm == 1)
        for c in range(C):
            assert np.array_equal(got[c], models[c].step(v[c], flush=fl, shift_mode=sm)), (L, i, c)
            if refs:
                assert np.array_equal(got[c], refs[c].step(v[c], flush=fl, shift_mode=sm)), (L, i, c, "compiled reference")
    assert refused or L == 2  # the scripts do reach the refused case


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_inputs_against_the_compiled_reference_itself(S, case):
    """No restatement in between: the golden cases' inputs go through oracle/_ref (the unmodified reference
    headers, compiled in the build container) on this box and through the CUDA path; outputs must be identical."""
    if O.ref() is None:
        pytest.skip("oracle/_ref did not travel to this box")
    import make_golden as MG
    assert np.array_equal(run_product(case), MG.run_reference(case))
