"""Shared drivers for the golden cases: run a case through the C oracle or through the product
classes (srcdsp_b200, CUDA behind the C ABI)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG  # noqa: E402  (parameters + tap designs only; no reference needed to import)

import oracle as O  # noqa: E402


def load_golden():
    z = np.load(os.path.join(HERE, "golden", "golden.npz"))
    cases = json.loads(bytes(z["cases"]).decode())
    return cases, z


def case_input(case):
    return O.corc().synth(case["seed"], 0, 0, case["n"], 0)


def run_oracle(case):
    """Plain-C restatement, streaming block by block with carried state."""
    c = O.corc()
    x = case_input(case)
    outs, pos, k = [], 0, case["kind"]
    if k in ("mixer", "mixer_adjust"):
        phi, fr, nominal = 0, c.mixer_set_frequency(case["f"]), case["f"]
        for i, b in enumerate(case["blocks"]):
            if k == "mixer_adjust" and i == 1:
                nominal = c.mixer_adjust_nominal(nominal, case["adj"])
                fr = c.mixer_set_frequency(nominal)
            y, phi = c.mixer_step(x[pos:pos + b], phi, fr)
            outs.append(y)
            pos += b
    elif k == "dec":
        h = None
        for b in case["blocks"]:
            y, h = c.dec_step(MG.taps_for(case), case["M"], x[pos:pos + b], h, case["left_shift"])
            outs.append(y)
            pos += b
    elif k == "fir":
        h = None
        for b in case["blocks"]:
            y, h = c.fir_step(MG.taps_for(case), x[pos:pos + b], h)
            outs.append(y)
            pos += b
    elif k == "ddc":
        phi, fr, h1, h2 = 0, c.mixer_set_frequency(case["f"]), None, None
        for b in case["blocks"]:
            y, phi = c.mixer_step(x[pos:pos + b], phi, fr)
            y, h1 = c.dec_step(MG.taps_for(case, "1"), case["M1"], y, h1)
            if case["M2"]:
                y, h2 = c.dec_step(MG.taps_for(case, "2"), case["M2"], y, h2)
            outs.append(y)
            pos += b
    elif k == "up":
        h = None
        for i, b in enumerate(case["blocks"]):
            last = i == len(case["blocks"]) - 1
            y, h = c.up_step(MG.taps_for(case), case["L"], x[pos:pos + b], h,
                             flush=last and case["flush_last"], shift_mode=case["shift_mode"])
            outs.append(y)
            pos += b
    return np.concatenate(outs)


def run_product(case, to_buf=lambda a: a, from_buf=lambda a: a, fused=True):
    """The CUDA path through srcdsp_b200 (C ABI).  to_buf/from_buf move blocks to the device for
    the device-pointer flavour (identity = host-pointer flavour)."""
    import srcdsp_b200 as S
    x = case_input(case)
    outs, pos, k = [], 0, case["kind"]
    if k in ("mixer", "mixer_adjust"):
        m = S.Mixer()
        m.setFrequency(case["f"])
        for i, b in enumerate(case["blocks"]):
            if k == "mixer_adjust" and i == 1:
                m.adjustFrequency(case["adj"])
            outs.append(from_buf(m.step(to_buf(x[pos:pos + b]))))
            pos += b
    elif k == "dec":
        d = S.FilterDnsamplingFir(case["M"], MG.taps_for(case), obsolete=True)
        d.setLeftShiftBy2(case["left_shift"])
        for b in case["blocks"]:
            outs.append(from_buf(d.step(to_buf(x[pos:pos + b]))))
            pos += b
    elif k == "fir":
        f = S.FilterFir(MG.taps_for(case))
        for b in case["blocks"]:
            outs.append(from_buf(f.step(to_buf(x[pos:pos + b]))))
            pos += b
    elif k == "ddc":
        m = S.Mixer()
        m.setFrequency(case["f"])
        d1 = S.FilterDnsamplingFir(case["M1"], MG.taps_for(case, "1"), obsolete=True)
        d2 = S.FilterDnsamplingFir(case["M2"], MG.taps_for(case, "2"), obsolete=True) if case["M2"] else None
        chain = S.Ddc(m, d1, d2) if fused else None
        for b in case["blocks"]:
            xb = to_buf(x[pos:pos + b])
            if fused:
                y = chain.step(xb)
            else:
                y = d1.step(m.step(xb))
                if d2:
                    y = d2.step(y)
            outs.append(from_buf(y))
            pos += b
    elif k == "up":
        u = S.FilterUpsamplingFir(case["L"], MG.taps_for(case))
        for i, b in enumerate(case["blocks"]):
            last = i == len(case["blocks"]) - 1
            outs.append(from_buf(u.step(to_buf(x[pos:pos + b]), flush=last and case["flush_last"],
                                        iterator_overload=case["shift_mode"] == 1)))
            pos += b
    return np.concatenate(outs)
