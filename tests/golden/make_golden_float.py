"""Generate tests/golden/decf.npz from the COMPILED REFERENCE (oracle/_ref): the float instantiation
dsptl::FilterDnsamplingFir<complex<float>, complex<float>, complex<float>, float, M> of the unmodified headers under
/root/reference.  Run in the build container only:

    python tests/golden/make_golden_float.py

Inputs are not stored: the counter-based synthetic baseband (oracle.orc_synth_fill) scaled to float by `scale`, and
deterministic taps, both re-derived by the tests from the parameters below.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

FCASES = [
    # unity-gain float taps (all |c| < 1: the reference's coeffScaling is its undefined-behaviour case, shift 0 on x86-64)
    dict(name="decf8_63_unity_cfg1", seed=0x5EED0F01, n=16384, blocks=[8192, 4096, 4096], M=8, ntaps=63, taps="unity", scale=0.37, left_shift=0),
    dict(name="decf16_255_unity_cfg2", seed=0x5EED0F02, n=32768, blocks=[16384, 16384], M=16, ntaps=255, taps="unity", scale=0.61, left_shift=0),
    dict(name="decf4_1023_unity_cfg5", seed=0x5EED0F05, n=16384, blocks=[8192, 8192], M=4, ntaps=1023, taps="unity", scale=0.5, left_shift=0),
    # integer-valued float taps (the integer filter's design): coeffScaling = 15
    dict(name="decf16_256_int", seed=0x5EED0F03, n=32768, blocks=[16384, 16384], M=16, ntaps=256, taps="int", scale=0.25, left_shift=0),
    dict(name="decf3_33_frac_ls1", seed=0x5EED0F06, n=6144, blocks=[3072, 1536, 1536], M=3, ntaps=33, taps="frac", scale=1.0, left_shift=1),
    dict(name="decf1_17_fir", seed=0x5EED0F07, n=4096, blocks=[1000, 3096], M=1, ntaps=17, taps="frac", scale=0.125, left_shift=0),
]


def ftaps_for(case) -> np.ndarray:
    nt, M = case["ntaps"], case["M"]
    if case["taps"] == "int":
        return O.design_lowpass_taps(nt, M).astype(np.float32)
    k = np.arange(nt, dtype=np.float64)
    h = np.hamming(nt) * np.sinc((k - (nt - 1) / 2.0) / max(M, 2))
    h = h / h.sum()
    if case["taps"] == "unity":
        return h.astype(np.float32)
    return (h * 97.3).astype(np.float32)  # "frac": non-integer taps with integer parts


def finput_for(case) -> np.ndarray:
    x = O.corc().synth(case["seed"], 0, 0, case["n"], 0).astype(np.float32)
    return (x * np.float32(case["scale"])).astype(np.float32)


def run_reference_float(case):
    r = O.ref()
    assert r is not None, "oracle/_ref is not built"
    f = O.RefDecF(r, case["M"], ftaps_for(case), obsolete=True)
    if case["left_shift"]:
        f.setLeftShiftBy2(case["left_shift"])
    x, outs, pos = finput_for(case), [], 0
    for n in case["blocks"]:
        outs.append(f.step(x[pos:pos + n]))
        pos += n
    return np.concatenate(outs)


if __name__ == "__main__":
    O.build()
    out = {c["name"]: run_reference_float(c) for c in FCASES}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decf.npz")
    np.savez_compressed(path, **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))
    print("wrote", path, os.path.getsize(path), "bytes")
