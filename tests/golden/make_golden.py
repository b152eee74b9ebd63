"""Generate tests/golden/golden.npz from the COMPILED REFERENCE (oracle/_ref, i.e. the unmodified
headers under /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are not stored: they are the counter-based synthetic baseband (oracle.orc_synth_fill,
full scale) and deterministic taps, re-derived by the tests from the parameters below.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

CASES = [
    # name, kind, params
    dict(name="mixer_f-0.3217", kind="mixer", seed=0x5EED0101, n=4096, blocks=[1000, 3096], f=-0.3217),
    dict(name="mixer_adjust", kind="mixer_adjust", seed=0x5EED0102, n=2048, blocks=[1024, 1024], f=0.25, adj=-0.4001),
    dict(name="dec8_63_cfg1", kind="dec", seed=0x5EED0001, n=16384, blocks=[8192, 4096, 4096], M=8, ntaps=63, left_shift=0),
    dict(name="dec16_255_cfg2", kind="dec", seed=0x5EED0002, n=32768, blocks=[16384, 16384], M=16, ntaps=255, left_shift=0),
    dict(name="dec4_1023_cfg5", kind="dec", seed=0x5EED0005, n=16384, blocks=[8192, 8192], M=4, ntaps=1023, left_shift=0),
    dict(name="dec3_31_ls1", kind="dec", seed=0x5EED0006, n=6144, blocks=[3072, 1536, 1536], M=3, ntaps=31, left_shift=1),
    dict(name="fir_33", kind="fir", seed=0x5EED0007, n=4096, blocks=[1000, 3096], ntaps=33),
    dict(name="ddc_mix_dec8_63", kind="ddc", seed=0x5EED0011, n=16384, blocks=[8192, 8192], f=-0.3217,
         M1=8, ntaps1=63, M2=0, ntaps2=0),
    dict(name="ddc_mix_8x4_cfg3", kind="ddc", seed=0x5EED0003, n=32768, blocks=[16384, 16384], f=0.1234,
         M1=8, ntaps1=63, M2=4, ntaps2=63),
    dict(name="up8_64_flush", kind="up", seed=0x5EED0004, n=2048, blocks=[1024, 1024], L=8, ntaps=64, shift_mode=0, flush_last=True),
    dict(name="up8_128", kind="up", seed=0x5EED0014, n=2048, blocks=[512, 1536], L=8, ntaps=128, shift_mode=0, flush_last=False),
    dict(name="up4_32_iter", kind="up", seed=0x5EED0024, n=1024, blocks=[1024], L=4, ntaps=32, shift_mode=1, flush_last=True),
]


def taps_for(case, which=""):
    if case["kind"] == "up":
        t = O.design_interp_taps(case["ntaps"], case["L"])
        if case["name"] == "up8_64_flush":
            t[-1] = 0  # exercises getLength() != getImpLength()
        if case["shift_mode"] == 1:
            t = (t // 8).astype(np.int32)  # iterator overload has shift 0: keep it out of saturation
        return t
    if case["kind"] == "fir":
        return O.design_lowpass_taps(case["ntaps"], 4)
    nt, M = case["ntaps" + which], case["M" + which]
    return O.design_lowpass_taps(nt, M)


def run_reference(case):
    r = O.ref()
    assert r is not None, "oracle/_ref is not built"
    x = O.corc().synth(case["seed"], 0, 0, case["n"], 0)
    outs, pos = [], 0
    k = case["kind"]
    if k in ("mixer", "mixer_adjust"):
        m = O.RefMixer(r)
        m.setFrequency(case["f"])
        for i, b in enumerate(case["blocks"]):
            if k == "mixer_adjust" and i == 1:
                m.adjustFrequency(case["adj"])
            outs.append(m.step(x[pos:pos + b]))
            pos += b
    elif k == "dec":
        d = O.RefDecimator(r, case["M"], taps_for(case), variant=0)
        d.setLeftShiftBy2(case["left_shift"])
        for b in case["blocks"]:
            outs.append(d.step(x[pos:pos + b]))
            pos += b
    elif k == "fir":
        f = O.RefFir(r, taps_for(case))
        for b in case["blocks"]:
            outs.append(f.step(x[pos:pos + b]))
            pos += b
    elif k == "ddc":
        m = O.RefMixer(r)
        m.setFrequency(case["f"])
        d1 = O.RefDecimator(r, case["M1"], taps_for(case, "1"), variant=0)
        d2 = O.RefDecimator(r, case["M2"], taps_for(case, "2"), variant=0) if case["M2"] else None
        for b in case["blocks"]:
            y = d1.step(m.step(x[pos:pos + b]))
            outs.append(d2.step(y) if d2 else y)
            pos += b
    elif k == "up":
        u = O.RefUpsampler(r, case["L"], taps_for(case))
        for i, b in enumerate(case["blocks"]):
            last = i == len(case["blocks"]) - 1
            outs.append(u.step(x[pos:pos + b], flush=last and case["flush_last"], shift_mode=case["shift_mode"]))
            pos += b
    return np.concatenate(outs)


if __name__ == "__main__":
    O.build()
    arrays = {"cases": np.frombuffer(json.dumps(CASES).encode(), dtype=np.uint8)}
    for c in CASES:
        arrays[c["name"]] = run_reference(c)
        print(c["name"], arrays[c["name"]].shape, int(np.abs(arrays[c["name"]].astype(np.int32)).max()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes; reference build:", O.ref().build_info())
    # FifoWithTimeTrack<cs16, 1024>: a scripted trace of operations and the reference's answers
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_fifo_files as T
    cap, seed = 1024, 11
    ops = T._script(seed, cap, 500)
    res = T._run(O.RefFifo(O.ref(), cap, 48000.0), ops, seed)
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fifo_trace.npz")
    np.savez_compressed(tp, capacity=cap, seed=seed, ops=np.array(ops, dtype=object), results=np.array(res, dtype=object))
    print("wrote", tp, os.path.getsize(tp), "bytes,", len(ops), "operations")
    # FixedPatternCorrelator: the reference's trace on two small cases
    import test_correlator as TC
    gold = {}
    for seed, N, S, n, blocks in [(21, 32, 4, 5000, [1200, 1800, 2000]), (22, 16, 2, 3000, [700, 300, 2000])]:
        pat, x, _ = TC.make_case(seed, N, S, n)
        rc = O.RefCorrelator(O.ref(), N, S)
        rc.setPattern(pat)
        gold[f"corr_{N}_{S}"] = np.array(dict(N=N, S=S, pattern=pat, x=x, blocks=blocks, trace=TC.run_blocks(rc, x, blocks)), dtype=object)
    cp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "correlator.npz")
    np.savez_compressed(cp, **gold)
    print("wrote", cp, os.path.getsize(cp), "bytes")
