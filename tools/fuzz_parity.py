"""Randomised differential test of the CUDA path against the C oracle (oracle/ is the checker here, as in tests/):

    python tools/fuzz_parity.py [seconds] [seed]

Random ratios, tap counts, tap magnitudes (1-3 byte digits and beyond), channel counts, streaming block lengths
(ragged, shorter than the filter, unaligned for the TMA feed), NCO frequencies and table sizes, left shifts, every
decimator kernel that accepts the shape (automatic, IMAD, band form, original TMA form), the upsampler (both shift
modes, flush) and the float decimator (pair and quad kernels).  Every comparison is array_equal.  Prints one line per
failure and a summary; exit code 1 if anything differed."""
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
import srcdsp_b200 as S  # noqa: E402

rng = None
corc = None
fails, cases = [], {"dec": 0, "ddc": 0, "up": 0, "decf": 0}
kernels = {}  # which kernels the passed cases ran (last step of the case)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rand_taps(nt):
    kind = rng.integers(5)
    amp = [3, 120, 30000, 2 ** 22, 2 ** 27][kind]  # 1, 1, 2, 3 digits, beyond the tensor-core path
    amp = max(1, min(amp, (2 ** 31 - 1) // nt))       # sum |c| < 2^31: the reference's shift count stays below 32
    t = rng.integers(-amp, amp + 1, nt).astype(np.int64)
    if rng.integers(3) == 0:  # sparse
        t[rng.random(nt) < 0.7] = 0
    if not t.any():
        t[nt // 2] = amp
    # keep the accumulator inside int32 for full-scale input only when the taps are small; otherwise it wraps, as in the reference
    return t.astype(np.int32)


def fuzz_dec(mix):
    M = int(rng.choice([1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32, 48, 64, 7, 9, 20]))
    nt = int(rng.integers(1, min(32 * M + 2, 1400) + 1)) if rng.integers(4) else int(rng.integers(1, 2100))
    taps = rand_taps(nt)
    C = int(rng.integers(1, 5))
    kind = int(rng.choice([0, 1, 2, 3, 4, 5, 4, 5]))  # (2 / 3 / 4 / 5: tcgen05 kernels, refused with -5 where they do not apply)
    ls = int(rng.integers(0, 3)) if nt > 4 else 0
    n_table = int(rng.choice([4096, 1024, 256, 8192])) if mix else 4096
    d = S.FilterDnsamplingFir(M, taps, channels=C, obsolete=True)
    d.setLeftShiftBy2(ls)
    chain, f = d, None
    if mix:
        m = S.Mixer(n_table, channels=C)
        f = rng.uniform(-1, 1, C).astype(np.float32)
        if rng.integers(2):
            f = (np.round(f * 64) / 64).astype(np.float32)  # short oscillator periods
        m.setFrequency(f)
        chain = S.Ddc(m, d)
    hs, ph = [None] * C, [0] * C
    align = M * 4 // math.gcd(M, 4)
    for blk in range(int(rng.integers(1, 5))):
        q = int(rng.choice([1, 7, 33, 70, 300, 2100, 9000, 40000]))
        n = M * q if rng.integers(3) == 0 else align * max(1, (M * q) // align)
        if kind in (4, 5):
            n = align * max(1, n // align)
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        d.set_kernel(kind)
        try:
            got = chain.step(dev(x)).cpu().numpy()
        except S.SrcDspError as e:
            if kind in (2, 3, 4, 5) and e.code == -5:  # the forced kernel does not take this shape: fall back to automatic
                d.set_kernel(0)
                kind = 0
                got = chain.step(dev(x)).cpu().numpy()
            else:
                raise
        for c in range(C):
            xi = x[c]
            if mix:
                fr = corc.mixer_set_frequency(float(f[c]), n_table)
                xi, ph[c] = corc.mixer_step(xi, ph[c], fr, n_table)
            e, hs[c] = corc.dec_step(taps, M, xi, hs[c], ls)
            if not np.array_equal(got[c], e):
                fails.append(("ddc" if mix else "dec", M, nt, C, kind, ls, n_table, blk, n, c, d.last_kernel))
                return
    cases["ddc" if mix else "dec"] += 1
    k = d.last_kernel.split(" ")[0]
    kernels[k] = kernels.get(k, 0) + 1


def fuzz_up():
    L = int(rng.choice([1, 2, 3, 4, 8, 16, 5, 32]))
    H = int(rng.integers(1, 20))
    nt = L * H
    amp = int(rng.choice([100, 4096, 30000, 2 ** 20]))
    taps = rng.integers(-amp, amp + 1, nt).astype(np.int32)
    if rng.integers(2):
        taps[-int(rng.integers(1, nt + 1)):] = 0  # trailing zeros: getLength() < getImpLength()
    if not taps.any():
        taps[0] = 1
    C = int(rng.integers(1, 4))
    u = S.FilterUpsamplingFir(L, taps, channels=C)
    hs = [None] * C
    for blk in range(int(rng.integers(1, 4))):
        n = int(rng.choice([1, 3, 64, 1000, 4099, 70000]))
        flush = bool(rng.integers(4) == 0)
        mode = int(rng.integers(2))
        x = rng.integers(-32768, 32768, (C, n, 2)).astype(np.int16)
        got = u.step(dev(x), flush=flush, iterator_overload=bool(mode)).cpu().numpy()
        for c in range(C):
            e, hs[c] = corc.up_step(taps, L, x[c], hs[c], flush, mode)
            if not np.array_equal(got[c], e):
                fails.append(("up", L, nt, C, mode, flush, blk, n, c, u.last_kernel))
                return
        if flush:
            break  # (the object's history after a flush is the zeros it pushed; stop the script here)
    cases["up"] += 1
    k = u.last_kernel.split(" ")[0]
    kernels[k] = kernels.get(k, 0) + 1


def fuzz_decf():
    M = int(rng.choice([1, 2, 3, 4, 8, 16, 32, 6, 12]))
    nt = int(rng.integers(1, 40 * M + 2))
    kind = int(rng.integers(3))
    if kind == 0:
        h = rng.normal(0, 1, nt)
        t = (h / max(1e-9, np.abs(h).sum())).astype(np.float32)
    elif kind == 1:
        t = rng.integers(-300, 301, nt).astype(np.float32)
    else:
        t = rng.uniform(-3, 3, nt).astype(np.float32)
    os.environ["SRCDSP_DECF_QUAD"] = str(int(rng.integers(2)))
    if rng.integers(3) == 0:
        os.environ["SRCDSP_DECF_CT"] = "0"
    else:
        os.environ.pop("SRCDSP_DECF_CT", None)
    C = int(rng.integers(1, 4))
    d = S.FilterDnsamplingFirFloat(M, t, channels=C, obsolete=True)
    hs = [None] * C
    for blk in range(int(rng.integers(1, 4))):
        n = M * int(rng.choice([1, 5, 130, 700, 5000]))
        x = rng.uniform(-30000, 30000, (C, n, 2)).astype(np.float32)
        got = d.step(dev(x)).cpu().numpy()
        for c in range(C):
            e, hs[c] = corc.decf_step(t, M, x[c], hs[c])
            if not np.array_equal(got[c], e):
                fails.append(("decf", M, nt, C, kind, os.environ["SRCDSP_DECF_QUAD"], blk, n, c, d.last_kernel))
                return
    cases["decf"] += 1
    k = d.last_kernel.split(" ")[0]
    kernels[k] = kernels.get(k, 0) + 1


def run(budget: float, seed: int, max_cases: int = 0):
    """Runs random cases for `budget` seconds (or until `max_cases` have been drawn: a deterministic set for a seed);
    returns (cases, failures, kernels)."""
    global rng, corc
    rng = np.random.default_rng(seed)
    O.build()
    corc = O.corc()
    fails.clear()
    kernels.clear()
    for k in cases:
        cases[k] = 0
    saved = {k: os.environ.get(k) for k in ("SRCDSP_DECF_QUAD", "SRCDSP_DECF_CT")}
    t0 = time.time()
    try:
        drawn = 0
        while time.time() - t0 < budget and len(fails) < 10 and (max_cases <= 0 or drawn < max_cases):
            drawn += 1
            r = rng.integers(10)
            if r < 4:
                fuzz_dec(False)
            elif r < 7:
                fuzz_dec(True)
            elif r < 8:
                fuzz_up()
            else:
                fuzz_decf()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return dict(cases), list(fails), dict(kernels)


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    t0 = time.time()
    c, f, k = run(budget, seed)
    for x in f:
        print("FAIL", x)
    print(f"fuzz_parity: seed {seed}, {time.time() - t0:.0f} s, passed cases {c}, failures {len(f)}")
    print("kernels of the passed cases:", dict(sorted(k.items())))
    sys.exit(1 if f else 0)
