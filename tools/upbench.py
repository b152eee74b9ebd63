"""Times FilterUpsamplingFir.step (device resident) for one shape with the CUDA-core kernels and with the tcgen05 kernel
(SRCDSP_UP_TC = 0 / 1, SRCDSP_UP_TC_FORM = 1 / 2):  python tools/upbench.py L ntaps [channels] [n_in]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srcdsp_b200 as S

L, nt = int(sys.argv[1]), int(sys.argv[2])
C = int(sys.argv[3]) if len(sys.argv) > 3 else 256
n = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
S.synth_fill(x, 0x5EED0004)
y = torch.empty((C, n * L, 2), dtype=torch.int16, device="cuda")
for tc, form in (("0", "1"), ("1", "1"), ("1", "2")):
    os.environ["SRCDSP_UP_TC"] = tc
    os.environ["SRCDSP_UP_TC_FORM"] = form
    u = S.FilterUpsamplingFir(L, S.design_interp_taps(nt, L), channels=C)
    for _ in range(3):
        u.step(x, out=y)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10):
        u.step(x, out=y)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    print(f"L={L} ntaps={nt} C={C} n={n} {u.last_kernel}: {ms:.3f} ms = {C * n * L / ms / 1e6:.0f} G out/s")
