"""srcdsp_b200.design (the benchmark's tap designer) == the test suite's twin in oracle/, and bench.py reaches into
oracle/ only where the contract lets it: the CPU-reference legs."""
import ast
import os

import numpy as np
import pytest

import oracle as O
import srcdsp_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ntaps,ratio", [(63, 8), (63, 4), (255, 16), (1023, 4), (31, 3), (33, 1), (129, 10), (511, 64)])
def test_lowpass_taps_equal_the_oracles(ntaps, ratio):
    a, b = S.design_lowpass_taps(ntaps, ratio), O.design_lowpass_taps(ntaps, ratio)
    assert a.dtype == np.int32 and np.array_equal(a, b)
    assert 32768 <= int(np.abs(a).sum()) <= 65535  # no int32 overflow for any int16 input, about unit gain after >> 15
    assert np.array_equal(a, a[::-1])              # linear phase
    pad = -(-ntaps // max(ratio, 1)) * max(ratio, 1)
    assert np.array_equal(S.design_lowpass_taps(ntaps, ratio, pad_to=pad), O.design_lowpass_taps(ntaps, ratio, pad_to=pad))


@pytest.mark.parametrize("ntaps,L", [(64, 8), (32, 4), (128, 8), (16, 2), (256, 16)])
def test_interp_taps_equal_the_oracles(ntaps, L):
    a, b = S.design_interp_taps(ntaps, L), O.design_interp_taps(ntaps, L)
    assert a.dtype == np.int32 and np.array_equal(a, b)
    for p in range(L):  # every polyphase branch carries about 32768 / L
        assert abs(int(a[p::L].sum()) - 32768 // L) <= 32768 // L // 8 + L


def test_lowpass_design_refuses_a_scale_outside_the_contract():
    with pytest.raises(ValueError):
        S.design_lowpass_taps(63, 8, target_sum=100000)


def test_bench_uses_the_oracle_only_in_its_cpu_legs():
    """`import oracle` may appear in bench.py only inside the functions that time the CPU reference (cpu_baseline /
    --impl reference) and, for the float workload, inside the `if not args.no_cpu` leg."""
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"cpu_cfg1_variants", "cpu_reference_run", "decf_bench"}
    found = set()

    def walk(node, fn):
        for ch in ast.iter_child_nodes(node):
            name = ch.name if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef)) and fn is None else fn
            if isinstance(ch, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in ch.names):
                found.add(fn)
            if isinstance(ch, ast.ImportFrom) and (ch.module or "").split(".")[0] == "oracle":
                found.add(fn)
            walk(ch, name)

    walk(tree, None)
    assert found <= allowed, f"bench.py imports oracle in {found - allowed}"
    assert {"cpu_cfg1_variants", "cpu_reference_run"} <= found
    # the float workload's import sits under its cpu_baseline guard
    src = open(os.path.join(ROOT, "bench.py")).read()
    i = src.index("def decf_bench")
    j = src.index("import oracle", i)
    assert "if not args.no_cpu:" in src[i:j] and src[i:j].rstrip().endswith("if not args.no_cpu:")
