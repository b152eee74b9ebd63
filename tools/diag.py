"""Runs one decimator step and, if the kernel trapped, asks the library which barrier wait timed out."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srcdsp_b200 as S
C, n, M, nt = 64, 1 << 22, 16, 255
x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
y = torch.empty((C, n // M, 2), dtype=torch.int16, device="cuda")
S.synth_fill(x, 1)
d = S.FilterDnsamplingFir(M, S.design_lowpass_taps(nt, M), channels=C, obsolete=True)
try:
    d.step(x, out=y)
    torch.cuda.synchronize()
    print("ok", d.last_kernel)
except Exception as e:
    print("failed:", str(e)[:100])
    try:
        d.step(x, out=y)
    except Exception as e2:
        print("diag:", e2)
