"""Times FilterDnsamplingFirFloat.step (device resident):  python tools/decfbench.py M ntaps [channels] [n_in]
Prints output rate, the FP32-pipe fraction (4 * ntaps rounded FP32 operations per output -- the reference's separate
multiply and add per component, which this kernel must keep -- against 148 SM x 128 lanes x SM clock) and the HBM
fraction (8 * M + 8 bytes per output against MEASURED_PEAKS.json)."""
import json
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
def run():
    d = S.FilterDnsamplingFirFloat(M, taps, channels=C, obsolete=True)
    for _ in range(3):
        d.step(x, out=y)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    K = 10
    for _ in range(K):
        d.step(x, out=y)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / K


# the pair kernel in its forced thread shapes, then the automatic choice (four outputs per thread where that applies)
# without L2 prefetch, the pair kernel without its whole-block fast path, and the library as shipped
for pairs, pf, blocks, quad in (("2", "1", "1", "0"), ("1", "1", "1", "0"), ("", "0", "1", ""), ("", "1", "0", "0"), ("", "1", "1", "")):
    for k in ("SRCDSP_DECF_PAIRS", "SRCDSP_DECF_QUAD"):
        os.environ.pop(k, None)
    if pairs:
        os.environ["SRCDSP_DECF_PAIRS"] = pairs
    if quad:
        os.environ["SRCDSP_DECF_QUAD"] = quad
    os.environ["SRCDSP_DECF_PREFETCH"] = pf
    os.environ["SRCDSP_DECF_BLOCKS"] = blocks
    ms = run()
    print(f"pairs per thread {pairs or 'auto'}, quad {quad or 'auto'}, L2 prefetch {pf}, whole blocks {blocks}: {ms:.3f} ms")
outs = C * (n // M)
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6549.8
fp32_peak = 148 * 128 * 1.965e9
print(f"M={M} ntaps={nt} C={C} n={n}: {ms:.3f} ms = {outs / ms / 1e6:.2f} G out/s; "
      f"FP32 pipe {outs * 4 * nt / (ms * 1e-3) / fp32_peak:.3f} of 148x128x1.965 GHz; "
      f"HBM {outs * (8 * M + 8) / (ms * 1e-3) / 1e9 / hbm:.3f} of {hbm} GB/s")
