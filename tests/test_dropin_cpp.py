"""The C++ drop-in headers (include/srcdsp/): a user program written against the reference API
compiles against both header sets; the reference build's output is the committed golden text and
the GPU build must print exactly the same."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "user_chain.cpp")
GOLDEN = os.path.join(ROOT, "tests", "golden", "user_chain.txt")
SRC_STREAM = os.path.join(ROOT, "tests", "cpp", "user_stream.cpp")       # buffers.h + dsptl_files.h
GOLDEN_STREAM = os.path.join(ROOT, "tests", "golden", "user_stream.txt")
REF = "/root/reference"
BUILD = os.path.join(ROOT, "build", "tests")


def _compile_dropin(out, src=SRC):
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(ROOT, "srcdsp_b200", "lib")
    cmd = ["g++", "-std=gnu++11", "-O1", "-I" + os.path.join(ROOT, "include", "srcdsp"), src, "-o", out,
           "-L" + lib, "-lsrcdsp_b200", "-Wl,-rpath," + lib]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_user_program_compiles_against_dropin_headers(built_lib):
    _compile_dropin(os.path.join(BUILD, "user_chain_gpu"))


def test_reference_build_matches_committed_golden():
    """Pins tests/golden/user_chain.txt to the unmodified reference (build container only)."""
    if not os.path.exists(os.path.join(REF, "mixers.h")):
        pytest.skip("no /root/reference here")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "user_chain_ref")
    subprocess.run(["g++", "-std=gnu++11", "-O2", "-w", "-I" + REF, SRC, os.path.join(REF, "dsp_complex.cpp"),
                    "-o", exe], check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    if os.environ.get("SRCDSP_REGEN_GOLDEN"):
        open(GOLDEN, "w").write(out)
    assert out == open(GOLDEN).read()


@pytest.mark.gpu
def test_dropin_build_prints_the_reference_output(built_lib):
    exe = os.path.join(BUILD, "user_chain_gpu")
    _compile_dropin(exe)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    assert out == open(GOLDEN).read()


def test_stream_program_reference_build_matches_committed_golden(tmp_path):
    """Pins tests/golden/user_stream.txt to the unmodified reference (build container only)."""
    if not os.path.exists(os.path.join(REF, "buffers.h")):
        pytest.skip("no /root/reference here")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "user_stream_ref")
    subprocess.run(["g++", "-std=gnu++11", "-O2", "-w", "-pthread", "-I" + REF, SRC_STREAM, "-o", exe], check=True,
                   capture_output=True, text=True)
    out = subprocess.run([exe, str(tmp_path / "ref.iq")], check=True, capture_output=True, text=True).stdout
    if os.environ.get("SRCDSP_REGEN_GOLDEN"):
        open(GOLDEN_STREAM, "w").write(out)
    assert out == open(GOLDEN_STREAM).read()


def test_stream_program_dropin_build_prints_the_reference_output(built_lib, tmp_path):
    """FifoWithTimeTrack + binary files through the drop-in headers: host logic, runs with or without a GPU
    (the ring is pinned when there is one)."""
    exe = os.path.join(BUILD, "user_stream_gpu")
    _compile_dropin(exe, SRC_STREAM)
    out = subprocess.run([exe, str(tmp_path / "gpu.iq")], check=True, capture_output=True, text=True, timeout=120).stdout
    assert out == open(GOLDEN_STREAM).read()


SRC_CORR = os.path.join(ROOT, "tests", "cpp", "user_corr.cpp")             # correlators.h
GOLDEN_CORR = os.path.join(ROOT, "tests", "golden", "user_corr.txt")


def test_corr_program_reference_build_matches_committed_golden():
    if not os.path.exists(os.path.join(REF, "correlators.h")):
        pytest.skip("no /root/reference here")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "user_corr_ref")
    subprocess.run(["g++", "-std=gnu++11", "-O2", "-w", "-I" + REF, SRC_CORR, os.path.join(REF, "dsp_complex.cpp"), "-o", exe],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    if os.environ.get("SRCDSP_REGEN_GOLDEN"):
        open(GOLDEN_CORR, "w").write(out)
    assert out == open(GOLDEN_CORR).read()
    assert "found 1" in out


@pytest.mark.gpu
def test_corr_program_dropin_build_prints_the_reference_output(built_lib):
    exe = os.path.join(BUILD, "user_corr_gpu")
    _compile_dropin(exe, SRC_CORR)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    assert out == open(GOLDEN_CORR).read()


SRC_FLOAT = os.path.join(ROOT, "tests", "cpp", "user_float.cpp")   # the float instantiation of the decimator
GOLDEN_FLOAT = os.path.join(ROOT, "tests", "golden", "user_float.txt")


def test_float_program_compiles_against_dropin_headers(built_lib):
    _compile_dropin(os.path.join(BUILD, "user_float_gpu"), SRC_FLOAT)


def test_float_program_reference_build_matches_committed_golden():
    """Pins tests/golden/user_float.txt (checksums over raw float bits) to the unmodified reference."""
    if not os.path.exists(os.path.join(REF, "dnsampling_filters.h")):
        pytest.skip("no /root/reference here")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "user_float_ref")
    subprocess.run(["g++", "-std=gnu++11", "-O2", "-w", "-I" + REF, SRC_FLOAT, os.path.join(REF, "dsp_complex.cpp"), "-o", exe],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    if os.environ.get("SRCDSP_REGEN_GOLDEN"):
        open(GOLDEN_FLOAT, "w").write(out)
    assert out == open(GOLDEN_FLOAT).read()


@pytest.mark.gpu
def test_float_program_dropin_build_prints_the_reference_output(built_lib):
    exe = os.path.join(BUILD, "user_float_gpu")
    _compile_dropin(exe, SRC_FLOAT)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    assert out == open(GOLDEN_FLOAT).read()


def test_mixer_base_class_is_exported(built_lib):
    """dsptl::_Mixer (reference mixers.h:26-41) exists with the same template parameters and public members: a class
    derived from it compiles against both header trees."""
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(BUILD, "user_mixer_base.cpp")
    open(src, "w").write(
        "#include <cassert>\n#include <cmath>\n#include <complex>\n#include <cstdint>\n#include <vector>\n#include \"mixers.h\"\n"
        "typedef std::complex<int16_t> cs16;\n"
        "struct Lo : dsptl::_Mixer<cs16, cs16, int16_t, 1024> {};\n"
        "int main() { Lo m; m.setFrequency(0.25f); m.adjustFrequency(-0.1f); m.reset(); m.reset(0.5f);\n"
        "  dsptl::Mixer<cs16, cs16, int16_t, 1024> x; dsptl::_Mixer<cs16, cs16, int16_t, 1024> &b = x; b.setFrequency(-0.5f); return 0; }\n")
    _compile_dropin(os.path.join(BUILD, "user_mixer_base_gpu"), src)
    if os.path.exists(os.path.join(REF, "mixers.h")):
        subprocess.run(["g++", "-std=gnu++11", "-O1", "-w", "-I" + REF, src, os.path.join(REF, "dsp_complex.cpp"), "-o",
                        os.path.join(BUILD, "user_mixer_base_ref")], check=True, capture_output=True, text=True)


@pytest.mark.parametrize("obsolete", [False, True])
def test_public_signatures_compile_against_both_trees(obsolete):
    """SURVEY.md 8(b) "signatures to keep" as static_asserts (tests/cpp/signatures.cpp: members, default arguments,
    return types, constness, what the obsolete header lacks): the same file must compile against the drop-in headers
    and -- where the reference checkout is present -- against the reference's own."""
    src = os.path.join(ROOT, "tests", "cpp", "signatures.cpp")
    trees = [os.path.join(ROOT, "include", "srcdsp")]
    if os.path.exists(os.path.join(REF, "mixers.h")):
        trees.append(REF)
    for inc in trees:
        cmd = ["g++", "-std=gnu++11", "-fsyntax-only", "-w", "-I" + inc, src]
        if obsolete:
            cmd.insert(3, "-DSRCDSP_SIGNATURES_OBSOLETE_HEADER")
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, f"{inc}:\n{r.stderr[-3000:]}"
