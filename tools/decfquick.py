"""Quick A/B of the float decimator under the environment's tuning variables:  python tools/decfquick.py M ntaps [C] [n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srcdsp_b200 as S

M, nt = int(sys.argv[1]), int(sys.argv[2])
C = int(sys.argv[3]) if len(sys.argv) > 3 else 256
n = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
h = np.hamming(nt) * np.sinc((np.arange(nt) - (nt - 1) / 2.0) / M)
taps = (h / h.sum()).astype(np.float32)
x = (torch.rand((C, n, 2), device="cuda") - 0.5) * 30000
y = torch.empty((C, n // M, 2), dtype=torch.float32, device="cuda")
d = S.FilterDnsamplingFirFloat(M, taps, channels=C, obsolete=True)
for _ in range(3):
    d.step(x, out=y)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record()
for _ in range(10):
    d.step(x, out=y)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
outs = C * (n // M)
env = " ".join(f"{k[7:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SRCDSP_DECF"))
print(f"M={M} ntaps={nt} [{env}] {d.last_kernel.split(' ')[0]}: {ms:.3f} ms, FP32 pipe {outs * 4 * nt / (ms * 1e-3) / (148 * 128 * 1.965e9):.3f}, checksum {float(y.double().sum()):.6e}")
