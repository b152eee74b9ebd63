"""include/srcdsp_b200.h is a C header: tests/c/abi_probe.c compiles as C99 with -pedantic -Werror, links against the
library and (a) on a box without a GPU sees every constructor refuse with SRCDSP_E_NOGPU, (b) on the B200 reproduces
the survey's decimator known-answer test through the ABI from plain C."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build", "tests")


def _build():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "abi_probe")
    lib = os.path.join(ROOT, "srcdsp_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_probe.c"), "-o", exe, "-L" + lib, "-lsrcdsp_b200", "-Wl,-rpath," + lib],
                   check=True, capture_output=True, text=True)
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_is_c99_and_constructors_refuse_without_a_gpu(built_lib):
    exe = _build()
    if _has_gpu():
        pytest.skip("this box has a GPU: the refusal path is the CPU box's test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "version 100"
    assert lines[1] == "device_count status -6 count 0"
    assert lines[2] == "mixer_create refused: -6 handle null message yes"
    assert lines[3] == "dec_create refused: -6 handle null" and lines[4] == "up_create refused: -6 handle null"
    assert lines[5] == "null handle: -1 0"  # a null handle is SRCDSP_E_INVALID; destroying one is a no-op


@pytest.mark.gpu
def test_known_answer_through_the_abi_from_plain_c(built_lib):
    r = subprocess.run([_build()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # 63 taps of 100 (sum 6300 -> shift 12), /8, constant (1000, -500) -> (1538, -770) after warm-up
    assert r.stdout.strip().splitlines()[-1] == "dec_step status 0 last (1538, -770)", r.stdout
