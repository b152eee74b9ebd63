"""A short run of tools/fuzz_parity.py (random decimator / fused-mixer / upsampler / float-decimator shapes, forced and
automatic kernels, array_equal against the oracle) inside the GPU suite: a fixed seed and 300 cases (~15 s).  The long runs are
recorded in profiles/r2_fuzz_parity.txt."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_shapes_against_the_oracle(built_lib):
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tools", "fuzz_parity.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    cases, fails, kernels = fz.run(300.0, 20261019, max_cases=300)
    assert not fails, fails
    assert sum(cases.values()) == 300, cases
    # the random kinds reach the tensor-core kernels, not only the IMAD fallback
    for k in ("dec_band_kernel", "dec_tma_kernel", "dec_tc_kernel", "dec_fir_kernel", "decf_quad_kernel", "decf_fir_kernel"):
        assert kernels.get(k, 0) > 0, kernels
