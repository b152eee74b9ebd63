"""Per-step kernel times of a workload over a long run (is the slowdown gradual = power / thermal, or constant?)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srcdsp_b200 as S
mixer = sys.argv[1] == "ddc16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
C, n, M, nt = 256, 1 << 24, 16, 255
x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
y = torch.empty((C, n // M, 2), dtype=torch.int16, device="cuda")
S.synth_fill(x, 0x5EED0002, amp_shift=2)
d = S.FilterDnsamplingFir(M, S.design_lowpass_taps(nt, M), channels=C, obsolete=True)
chain = d
if mixer:
    m = S.Mixer(channels=C)
    m.setFrequency((-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32))
    chain = S.Ddc(m, d)
torch.cuda.synchronize()
time.sleep(1.0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for i in range(steps):
    chain.step(x, out=y)
    ev[i + 1].record()
torch.cuda.synchronize()
t = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
print(sys.argv[1], " ".join(f"{v:.2f}" for v in t))
