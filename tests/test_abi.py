"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/srcdsp_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "srcdsp_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srcdsp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_api():
    names = header_functions()
    # one entry point per reference member on the hot path
    for n in ["srcdsp_mixer_create", "srcdsp_mixer_set_frequency", "srcdsp_mixer_reset",
              "srcdsp_mixer_adjust_frequency", "srcdsp_mixer_step",
              "srcdsp_dec_create", "srcdsp_dec_set_coeffs", "srcdsp_dec_step", "srcdsp_dec_reset",
              "srcdsp_dec_set_left_shift",
              "srcdsp_up_create", "srcdsp_up_set_coefficients", "srcdsp_up_step", "srcdsp_up_reset",
              "srcdsp_up_get_length", "srcdsp_up_get_imp_length", "srcdsp_up_get_ratio",
              "srcdsp_ddc_create", "srcdsp_ddc_step",
              # the float instantiation of the decimator (and of FilterFir = its M = 1 case)
              "srcdsp_decf_create", "srcdsp_decf_set_coeffs", "srcdsp_decf_step", "srcdsp_decf_reset",
              "srcdsp_decf_set_left_shift"]:
        assert n in names


def test_library_exports_every_declared_symbol(built_lib):
    from srcdsp_b200 import _capi
    names = header_functions()
    assert names, "no functions parsed from the header"
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in the header but not exported"
    assert sorted(_capi.PROTOTYPES) == names, "ctypes prototype table out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (srcdsp_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_library_is_sm100a_only(built_lib):
    from srcdsp_b200 import _capi
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    out = subprocess.run([cuobjdump, "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from srcdsp_b200 import _capi
    h = C.c_void_p()
    st = built_lib.srcdsp_dec_create(C.byref(h), 0, 1, 8)
    assert st == _capi.E_NOGPU and not h
    assert b"no CPU fallback" in built_lib.srcdsp_last_error()
    import srcdsp_b200 as S
    with pytest.raises(S.SrcDspError):
        S.Mixer()
    # the multi-device driver refuses as well (and cleans up after the member that failed)
    g, devs = C.c_void_p(), (C.c_int * 2)(0, 1)
    st = built_lib.srcdsp_group_create(C.byref(g), 1, devs, 2, 1, 4096, 8, 4)
    assert st == _capi.E_NOGPU and not g
    with pytest.raises(S.SrcDspError):
        S.DdcGroup("channels", [0], 4, 8, [1] * 8)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (checked statically)."""
    pkg = os.path.join(ROOT, "srcdsp_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                s = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", s, flags=re.M), f
                assert "liborc" not in s and "srcdsp_oracle.h" not in s, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().replace("oracle/srcdsp_oracle.c:orc_synth_fill", "")


def test_no_exception_crosses_the_abi():
    """never throw across the ABI.  Every extern "C" entry point with a body of its own is a
    function-try-block that ends in SRCDSP_ABI_CATCH (checked statically: provoking std::bad_alloc is not a unit test)."""
    csrc = os.path.join(ROOT, "srcdsp_b200", "csrc")
    guarded = 0
    for f in ("capi.cu", "capi_decf.inc", "fifo.cu", "group.cu"):
        lines = open(os.path.join(csrc, f)).read().split("\n")
        lo = next(i for i, l in enumerate(lines) if l.startswith('extern "C" {'))
        hi = next(i for i, l in enumerate(lines) if l.startswith('}  // extern "C"'))
        i = lo
        while i < hi:
            l = lines[i]
            if re.match(r"^\w[\w \*]*\bsrcdsp_\w+\(", l) and not l.startswith("static") and not l.rstrip().endswith(";"):
                if l.rstrip().endswith("}"):  # one-line accessor: returns a field, allocates nothing
                    assert "new " not in l and "std::" not in l, (f, l)
                    i += 1
                    continue
                j = i
                while lines[j].strip() not in ("{", "try {"):
                    j += 1
                assert lines[j] == "try {", f"{f}:{i + 1}: {l.strip()} is not a function-try-block"
                k = j + 1
                while lines[k] != "}":
                    k += 1
                assert lines[k + 1] == "SRCDSP_ABI_CATCH", f"{f}:{k + 2}"
                guarded += 1
                i = k + 1
            else:
                i += 1
    assert guarded >= 80
