"""Band form (dec_band_kernel, kind 4) against the original form (dec_tma_kernel, kind 5) of the tcgen05 decimator over
ratios / filter lengths, with and without the fused mixer: 256 channels x 8 Mi samples, burst timing (8 steps after a
0.5 s pause).  Decides the automatic choice in DecBank::step_device."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srcdsp_b200 as S


def run(M, nt, C, n, kind, mix):
    n = n // (32 * M) * (32 * M)
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, 1, amp_shift=2)
    y = torch.empty((C, n // M, 2), dtype=torch.int16, device="cuda")
    d = S.FilterDnsamplingFir(M, S.design_lowpass_taps(nt, M), channels=C, obsolete=True)
    d.set_kernel(kind)
    ch = d
    if mix:
        m = S.Mixer(channels=C)
        m.setFrequency((-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32))
        ch = S.Ddc(m, d)
    for _ in range(3):
        ch.step(x, out=y)
    torch.cuda.synchronize()
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        ch.step(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 8, d.last_kernel[:8]


for M, nt in [(2, 33), (4, 31), (4, 127), (6, 97), (8, 63), (8, 255), (10, 160), (12, 255), (16, 127), (16, 255), (16, 511), (32, 255), (32, 1023), (64, 511)]:
    for mix in (False, True):
        a = run(M, nt, 256, 1 << 23, 5, mix)
        b = run(M, nt, 256, 1 << 23, 4, mix)
        print(f"M={M:2d} nt={nt:4d} mix={int(mix)}: original {a[0]:.3f} ms ({a[1]})  band {b[0]:.3f} ms ({b[1]})  band/original {b[0] / a[0]:.3f}", flush=True)
