#!/usr/bin/env python
"""bench.py -- headline benchmark of the SrcDsp DDC hot path on B200 (contract: see README/DESIGN).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

Workload (BASELINE.json configs[1], "cfg2"): decimate-by-16, 255-tap polyphase FIR over 256
independent complex int16 channels x 16 Mi samples each, per GPU (weak scaling: N GPUs run
N x 256 channels, contiguous channel batches per rank, no data-path collective).

One "step" = one pass of the decimator bank over the whole 16 GiB device-resident batch
(much larger than the 126 MB L2, so no L2 flush is needed between timed iterations).

  value      whole-job output Msamples/s, input already resident in HBM, CUDA-event timed
  e2e        same metric through the public API with HOST (pinned) buffers: H2D + kernels + D2H
  roofline   dominant kernel's algorithmic bytes / its measured duration vs measured HBM peak
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) timed on this box's host cores

`--impl reference` times the reference's own CPU implementation (oracle/_ref, all host threads)
on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x5EED0002
WORKLOADS = {
    # kind: dec = decimator, ddc = NCO mix + decimator, ddc2 = mix + two decimators, up = interpolator
    "cfg2": dict(kind="dec", channels=256, n=1 << 24, M=16, ntaps=255, mix=False,
                 desc="cfg2: decimate-by-16 255-tap polyphase FIR, 256 ch x 16Mi cs16 samples per GPU"),
    "ddc16": dict(kind="ddc", channels=256, n=1 << 24, M=16, ntaps=255, mix=True,
                  desc="ddc16: NCO mix fused into decimate-by-16 255-tap FIR, 256 ch x 16Mi per GPU"),
    "ddc8": dict(kind="ddc", channels=1024, n=1 << 22, M=8, ntaps=63, mix=True,
                 desc="ddc8: NCO mix fused into decimate-by-8 63-tap FIR, 1024 ch x 4Mi per GPU (stage 1 of cfg3)"),
    "cfg1": dict(kind="ddc", channels=1, n=1 << 20, M=8, ntaps=63, mix=True,
                 desc="cfg1: NCO mix + decimate-by-8 63-tap FIR, 1 ch x 1Mi samples"),
    "cfg3": dict(kind="ddc2", channels=1024, n=1 << 22, M=8, ntaps=63, M2=4, ntaps2=63, mix=True,
                 desc="cfg3: NCO mix + /8 (63 taps) + /4 (63 taps), 1024 ch x 4Mi samples (per GPU when weak scaling)"),
    "cfg4": dict(kind="up", channels=256, n=1 << 20, M=8, ntaps=64, mix=False,
                 desc="cfg4: interpolate-by-8 64-tap FIR, 256 ch x 1Mi input samples (8Mi out) per step"),
    "cfg5": dict(kind="dec", channels=1, n=1 << 29, M=4, ntaps=1023, mix=False,
                 desc="cfg5 slice: decimate-by-4 1023-tap FIR, one stream slice of 512Mi samples per GPU"),
    "mid": dict(kind="dec", channels=64, n=1 << 22, M=16, ntaps=255, mix=False, desc="mid: 64 ch x 4Mi, /16, 255 taps"),
    "mix": dict(kind="mix", channels=256, n=1 << 24, M=1, ntaps=0, mix=True,
                desc="mix: stand-alone NCO mixer (mixers.h Mixer::step), 256 ch x 16Mi samples"),
    "corr": dict(kind="corr", channels=256, n=1 << 22, M=1, ntaps=32, mix=False, S=4,
                 desc="corr: FixedPatternCorrelator<int16,int32,32,4> bank, 256 ch x 4Mi samples of noise (no peak: full scan)"),
    "fifo": dict(kind="fifo", channels=1, n=1 << 22, M=16, ntaps=255, mix=False, blocks=24,
                 desc="fifo: one stream through FifoWithTimeTrack (pinned ring, 16Mi samples) -> decimate-by-16 255-tap FIR, 4Mi-sample blocks"),
    "cfg2f": dict(kind="decf", channels=256, n=1 << 22, M=16, ntaps=255, mix=False,
                  desc="cfg2f: the reference's FLOAT instantiation of cfg2's filter (complex<float> samples, float taps), "
                       "decimate-by-16 255-tap FIR, 256 ch x 4Mi samples (8 GiB in) per GPU"),
    "cfg1f": dict(kind="decf", channels=256, n=1 << 22, M=8, ntaps=63, mix=False,
                  desc="cfg1f: float decimate-by-8 63-tap FIR (BASELINE configs[0]'s filter), 256 ch x 4Mi samples per GPU"),
    "smoke": dict(kind="dec", channels=8, n=1 << 18, M=16, ntaps=255, mix=False, desc="smoke: 8 ch x 256Ki"),
}


def bytes_per_out(w):
    """Algorithmic HBM bytes per output sample: cs16 in + cs16 out (+ the stage-1
    output written and re-read for the two-stage chain)."""
    if w["kind"] == "up":
        return 4.0 + 4.0 / w["M"]
    if w["kind"] == "mix":
        return 8.0
    if w["kind"] == "ddc2":
        return 4.0 * w["M"] * w["M2"] + 4.0  # 132 B; the stage-1 round trip is traffic, not algorithm
    return 4.0 * w["M"] + 4.0


def traffic_model_bytes_per_out(w):
    """What the kernels really move per output: the two-stage chain writes and re-reads the stage-1 stream."""
    return bytes_per_out(w) + (8.0 * w["M2"] if w["kind"] == "ddc2" else 0.0)


BINDING = {  # which resource bounds the dominant kernel of a workload; the roofline fraction is always vs HBM
    "cfg2": "hbm (sustained: the boxes' 1 kW power cap lowers the SM clock first)",
    "ddc16": "sm issue / integer pipes of the fused mixer, then hbm",
    "ddc8": "hbm", "cfg3": "hbm (+ the stage-1 round trip)", "cfg4": "int32 multiply pipe (16 IMAD per output), hbm writes next",
    "cfg5": "tensor pipe + shared-memory operand reads (2046 MACs per output)", "cfg1": "launch latency (131072 outputs)",
    "mix": "hbm", "mid": "hbm",
}


def out_per_in(w):
    if w["kind"] == "up":
        return float(w["M"])
    return 1.0 / (w["M"] * w.get("M2", 1))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polling threads (clock + reasons about
    every 2 ms, board power on its own thread; the timed region of the default run is ~30-60 ms, too short for an
    `nvidia-smi -lms` child to report anything but the idle GPU afterwards), `nvidia-smi` only as a fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.samples, self.stop_flag, self.t = None, None, [], False, None
        self.power, self.tp = [], None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                for cand in (uuid if uuid.startswith("GPU-") else "GPU-" + uuid,):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if hasattr(cand, "encode") else cand)
                    except Exception:
                        try:
                            h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                        except Exception:
                            h = None
            except Exception:
                h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.handle = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        except Exception:
            self.nvml = None

    def _poll(self):
        # clock + throttle reasons: two cheap queries per sample
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)), int(self._reasons(self.handle))))
            except Exception:
                pass
            time.sleep(0.002)

    def _poll_power(self):
        # board power on its own thread: the query can take 10-20 ms, which left the headline's 30 ms region with a
        # single clock sample when it shared the loop above
        n = self.nvml
        while not self.stop_flag:
            try:
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            self.samples, self.power, self.stop_flag = [], [], False
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.tp = threading.Thread(target=self._poll_power, daemon=True)
            self.t.start()
            self.tp.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            for th in (self.t, self.tp):
                if th:
                    th.join(timeout=2)
            sm = [s[0] for s in self.samples]
            mask = 0
            for s in self.samples:
                mask |= s[1]
            reasons = sorted(nm for bit, nm in self.REASONS.items() if mask & bit)
            pw = list(self.power)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                    "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                    "power_w": float(np.median(pw)) if pw else None, "power_w_max": float(max(pw)) if pw else None,
                    "source": "nvml, sampled inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ----------------------------------------------------------------------------------------------
def cpu_cfg1_variants(w):
    """SURVEY.md 8(d) CPU baselines (A) and (B) on BASELINE configs[0] EXACTLY (one channel, NCO mix + decimate-by-8
    63-tap FIR over 1 Mi samples, one step() call per stage): the unmodified reference built with its makefile's flags
    (-std=gnu++11, no -O; makefile:18) and with -O2, one host thread each.  Reported, never a target."""
    import oracle as O
    c = O.corc()
    n, M, nt = w["n"], w["M"], w["ntaps"]
    x = c.synth(SEED, 0, 0, n, 2)[None]
    taps = O.design_lowpass_taps(nt, M)
    lo = np.array([-0.3217], np.float32)
    out = {}
    for name, opt, flags in (("makefile_build", "O0", "-std=gnu++11 (reference makefile:18, no -O)"), ("O2", "O2", "-std=gnu++11 -O2")):
        r = O.ref(opt)
        if r is None:
            out[name] = {"value": None, "error": f"oracle/_ref ({opt}) is not built"}
            continue
        r.bench_bank(1, x, n, 1, M, taps, lo_freq=lo)  # warm-up
        secs = [r.bench_bank(1, x, n, 1, M, taps, lo_freq=lo)[0] for _ in range(3)]
        t = float(np.median(secs))
        out[name] = {"value": (n // M) / t / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "reference", "seconds": t,
                     "sample": f"cfg1 exactly: 1 channel x {n} samples, mixer.step + decimator.step, g++ {flags}, median of 3"}
    return out


def cpu_reference_run(w, steps: int, warmup: int, cores: int, target_cpu_seconds: float = 20.0):
    """Times the unmodified reference (oracle/_ref) on a bounded sample of the workload."""
    import oracle as O
    r = O.ref()
    kind = "reference"
    if r is None:
        raise RuntimeError("oracle/_ref is not built (run `make -C oracle` where /root/reference exists)")
    if w["kind"] not in ("dec", "ddc", "ddc2", "up"):
        raise RuntimeError(f"the reference harness has no bank benchmark for workload kind '{w['kind']}'")
    M, nt = w["M"], w["ntaps"]
    up = w["kind"] == "up"
    taps = O.design_interp_taps(nt, M) if up else O.design_lowpass_taps(nt, M)
    taps2 = O.design_lowpass_taps(w["ntaps2"], w["M2"]) if w["kind"] == "ddc2" else None
    ch = min(w["channels"], 2 * cores)
    block = 1 << 16
    c = O.corc()
    lo = (-1 + 2 * (np.arange(ch) + 0.5) / ch).astype(np.float32) if w["mix"] else None
    wk = {"dec": 0, "ddc": 1, "ddc2": 2, "up": 3}[w["kind"]]  # harness selector (not the cpu_baseline "kind")
    # pilot run on one 64Ki block per channel sizes the sample for ~target_cpu_seconds of CPU work
    xp = np.stack([c.synth(SEED, k, 0, block, 2) for k in range(ch)])
    tp, _ = r.bench_bank(wk, xp, block, cores, M, taps, w.get("M2", 1), taps2, lo_freq=lo)
    wall = target_cpu_seconds / cores
    n = int(block * max(1.0, wall / max(tp, 1e-6))) // block * block
    n = max(block, min(n, w["n"], (2 << 30) // (4 * ch) // block * block))
    x = np.stack([c.synth(SEED, k, 0, n, 2) for k in range(ch)])
    times = []
    for i in range(warmup + steps):
        secs, _ = r.bench_bank(wk, x, block, cores, M, taps, w.get("M2", 1), taps2, lo_freq=lo)
        if i >= warmup:
            times.append(secs)
    n_out = int(ch * n * out_per_in(w))
    t = float(np.mean(times))
    return dict(value=n_out / t / 1e6, unit="Msamples/s", cores=cores, kind=kind,
                sample=f"{ch} channels x {n} input samples of {w['desc'].split(':')[0]}, {r.build_info()}, "
                       f"{cores} threads, 64Ki-sample streaming blocks, mean of {len(times)} runs",
                seconds=t), t


def fifo_stream_bench(args, w, base, S, torch, device):
    """a producer thread writes time-stamped blocks into the pinned FifoWithTimeTrack, the
    consumer takes zero-copy segments of the ring and runs the decimator on them (H2D DMA straight from the ring,
    D2H of the outputs).  End-to-end by construction: value == e2e."""
    n, M, nt, blocks = w["n"], w["M"], w["ntaps"], w["blocks"]
    fifo = S.FifoWithTimeTrack(4 * n, 1e8)
    dec = S.FilterDnsamplingFir(M, S.design_lowpass_taps(nt, M), channels=1, device=device, obsolete=True)
    src = S.PinnedBuffer(1, n)
    blk = torch.empty((1, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(blk, SEED, ch0=0, n0=0, amp_shift=2)
    src.array[0] = blk[0].cpu().numpy()
    del blk
    out = S.PinnedBuffer(1, n // M)
    total = args.warmup + args.steps * blocks
    state = {"written": 0, "stop": False}

    def producer():
        for b in range(total):
            while state["written"] - state.get("read", 0) >= 3 and not state["stop"]:
                time.sleep(0)
            fifo.write(src.array[0], seconds=b, fracSeconds=0.0)
            state["written"] = b + 1

    th = threading.Thread(target=producer, daemon=True)
    th.start()
    start, done, t0 = 1, 0, None
    sampler = ClockSampler(device)
    for b in range(total):
        if b == args.warmup:
            torch.cuda.synchronize()
            sampler.start()
            l0 = S.launch_count()
            t0 = time.perf_counter()
        while state["written"] <= b:
            time.sleep(0)
        err, st, segs = fifo.readSegments(n, start)
        assert not err and st == start, (err, st, start)
        pos = 0
        for sgm in segs:  # one piece, or two when the block wraps around the end of the ring (multiples of M here)
            dec.step(sgm, out=out.array[0, pos // M: (pos + sgm.shape[0]) // M])
            pos += sgm.shape[0]
        start += n
        state["read"] = b + 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = S.launch_count() - l0
    state["stop"] = True
    th.join()
    n_out = args.steps * blocks * (n // M)
    value = n_out / dt / 1e6
    line = dict(base, value=value, ms_per_step=dt / args.steps * 1e3, roofline=None, cpu_baseline=None,
                e2e={"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": blocks * n * 4, "d2h_bytes_per_step": blocks * n // M * 4,
                     "api": "FifoWithTimeTrack.write (producer thread) -> readSegments -> FilterDnsamplingFir.step(pinned ring views)"},
                clocks=clocks, gpu_launches=int(launches), impl="ours")
    line["config"]["note"] = (f"one step = {blocks} blocks of {n} samples; host memcpy into the ring + PCIe bound, "
                              f"ring pinned: {fifo.pinned}")
    print(json.dumps(line))
    return 0


def corr_bench(args, w, base, S, torch, device):
    """the correlator bank scanning device-resident noise (no detection, so every sample is
    visited).  Metric here is INPUT Msamples/s; algorithmic bytes = 4 B per input sample (the scan only reads)."""
    C, n, N, St = w["channels"], w["n"], w["ntaps"], w["S"]
    x = torch.empty((C, n, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, SEED, amp_shift=6)  # +-512: stays below the correlator's energy > 300 / 2.7x threshold
    rng = np.random.default_rng(7)
    corr = S.FixedPatternCorrelator(N, St, channels=C, device=device)
    corr.setPattern(((rng.integers(0, 2, (N, 2)) * 2 - 1) * 1500).astype(np.int32))
    for _ in range(args.warmup):
        f, _ = corr.step(x)
    assert not f.any(), "the benchmark input must not contain a peak"
    torch.cuda.synchronize()
    sampler = ClockSampler(device)
    sampler.start()
    l0 = S.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    evs[0].record()
    for _ in range(args.steps):
        corr.step(x)
    evs[1].record()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[1]) / args.steps
    clocks = sampler.stop()
    peak, peak_src = peaks()
    alg = 4.0 * C * n
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    traffic = json.load(open(tp)).get("corr") if os.path.exists(tp) else None
    line = dict(base, metric="input Msamples/s", value=C * n / (ms * 1e-3) / 1e6, ms_per_step=ms,
                roofline={"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": traffic, "kernel": "corr_scan_blocked_kernel",
                          "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": ms,
                          "note": "step() is synchronous (it returns found / corrIndex): ms includes the 1 KB D2H of the result; "
                                  "2*N = 64 complex multiply-adds + 64 energy terms per sample put the scan on the IMAD pipe"},
                cpu_baseline=None, e2e=None, clocks=clocks, gpu_launches=int(S.launch_count() - l0), impl="ours")
    print(json.dumps(line))
    return 0


def decf_bench(args, w, base, S, torch, device):
    """The float instantiation FilterDnsamplingFir<complex<float>, ..., float, M> (bit-exact with the reference's
    tap-order float sum).  Algorithmic bytes per output: 8*M + 8; arithmetic: 4*ntaps rounded FP32 operations per
    output (a separate multiply and add per component and tap -- fusing them would change the result), which makes the
    FP32 pipe, not HBM, the binding roofline for cfg2's filter."""
    C, n, M, nt = w["channels"], w["n"], w["M"], w["ntaps"]
    h = np.hamming(nt) * np.sinc((np.arange(nt) - (nt - 1) / 2.0) / M)
    taps = (h / h.sum()).astype(np.float32)
    x = (torch.rand((C, n, 2), device="cuda", generator=torch.Generator(device="cuda").manual_seed(SEED)) - 0.5) * 16384
    y = torch.empty((C, n // M, 2), dtype=torch.float32, device="cuda")
    d = S.FilterDnsamplingFirFloat(M, taps, channels=C, device=device, obsolete=True)
    for _ in range(args.warmup):
        d.step(x, out=y)
    torch.cuda.synchronize()
    sampler = ClockSampler(device)
    sampler.start()
    l0 = S.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    evs[0].record()
    for _ in range(args.steps):
        d.step(x, out=y)
    evs[1].record()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[1]) / args.steps
    clocks = sampler.stop()
    launches = int(S.launch_count() - l0)
    peak, peak_src = peaks()
    n_out = C * (n // M)
    alg = (8.0 * M + 8.0) * n_out
    roof = {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
            "traffic": None, "kernel": d.last_kernel, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg, "kernel_ms": ms}
    if clocks and clocks.get("sm_max_mhz"):
        roof["fp32_frac"] = 4.0 * nt * n_out / (ms * 1e-3) / (148 * 128 * clocks["sm_max_mhz"] * 1e6)
        roof["fp32_note"] = ("4*ntaps rounded FP32 operations per output against 148 SM x 128 lanes x sm_max_mhz; the reference's "
                             "float sum is a multiply and an add per tap and component (no FMA), kept for bit-exactness; the kernel issues them "
                             "as packed FFMA2(x, k, -0.0) + FADD2 on (re, im): the same two roundings per component")
    cpu = None
    if not args.no_cpu:
        import oracle as O  # cpu_baseline leg: the compiled reference
        r = O.ref()
        if r is not None:  # the unmodified reference's float instantiation, one thread, a bounded sample of the workload
            import time
            ns = 1 << 20
            xs = x[0, :ns].cpu().numpy()
            f = O.RefDecF(r, M, taps, obsolete=True)
            f.step(xs[: 1 << 16])
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 10.0:
                f.step(xs)
                reps += 1
            dt = time.perf_counter() - t0
            cpu = {"value": reps * (ns // M) / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "reference",
                   "sample": f"{reps} x 1 channel x {ns} input samples of {args.workload}, SrcDsp reference headers (float instantiation), "
                             "unmodified, g++ optimised, 1 thread"}
    # end to end: host float buffers through the C ABI (pinned), H2D + kernel + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        import time
        Ce = min(C, 64)
        hx, hy = S.PinnedBuffer(Ce, 2 * n), S.PinnedBuffer(Ce, 2 * (n // M))  # 8 bytes per complex float sample
        xin = hx.array.view(np.float32).reshape(Ce, n, 2)
        yout = hy.array.view(np.float32).reshape(Ce, n // M, 2)
        xin[...] = x[:Ce].cpu().numpy()
        de = S.FilterDnsamplingFirFloat(M, taps, channels=Ce, device=device, obsolete=True)
        de.step(xin, out=yout)
        t0 = time.perf_counter()
        for _ in range(2):
            de.step(xin, out=yout)
        dt = (time.perf_counter() - t0) / 2
        e2e = {"value": Ce * (n // M) / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(xin.nbytes),
               "d2h_bytes_per_step": int(yout.nbytes), "steps": 2, "ms_per_step": dt * 1e3,
               "api": f"FilterDnsamplingFirFloat.step(host numpy float32, {Ce} channels) -> C ABI srcdsp_decf_step",
               "pinned": True}
    line = dict(base, dtype="f32", value=n_out / (ms * 1e-3) / 1e6, ms_per_step=ms, roofline=roof, cpu_baseline=cpu, e2e=e2e,
                clocks=clocks, gpu_launches=launches, impl="ours")
    print(json.dumps(line))
    return 0


def h2d_ceiling(torch, S, device, dist, barrier, all_max, chunk_bytes=48 << 20, total_bytes=4 << 30):
    """The box's ceiling for the end-to-end number: every rank copies pinned host memory to its GPU with one bare
    cudaMemcpyAsync per 48 MB chunk (nothing batched, no kernels), all ranks at once; D2H the same way afterwards."""
    src = torch.empty(chunk_bytes * 4, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(chunk_bytes * 4, dtype=torch.uint8, device="cuda")
    out = {}
    for name, (a, b) in (("h2d", (dst, src)), ("d2h", (src, dst))):
        n_chunks = total_bytes // chunk_bytes
        barrier()
        t0 = time.perf_counter()
        for i in range(n_chunks):
            o = (i % 4) * chunk_bytes
            a[o:o + chunk_bytes].copy_(b[o:o + chunk_bytes], non_blocking=True)
        torch.cuda.synchronize()
        dt_own = time.perf_counter() - t0
        barrier()
        dt = all_max(dt_own)
        out[name + "_gbs_per_gpu_slowest"] = n_chunks * chunk_bytes / dt / 1e9
        out[name + "_gbs_this_rank"] = n_chunks * chunk_bytes / dt_own / 1e9
    out["chunk_bytes"] = chunk_bytes
    out["note"] = "bare cudaMemcpyAsync from / to pinned memory, one call per chunk, all ranks concurrently; per-GPU rate of the slowest rank"
    return out


STREAM_SAMPLES = 1 << 32   # BASELINE configs[4]: "single very long stream (4G complex samples)"
STREAM_CALL = 1 << 29      # samples per step() call (the reference indexes with int: < 2^31 per call)


def stream_bench(args, w, base, S, torch, dist, rank, world, local_rank, timed, roofline_of, time_slices):
    """BASELINE configs[4] as a fixed job (strong scaling): ONE stream of 2^32 complex samples, decimate-by-4 1023-tap FIR,
    cut into `world` time slices at multiples of the decimation.  Rank g holds samples [start - warm, start + length) of
    the stream in HBM; a step = reset, warm-up run over the halo in front of the slice (outputs discarded; rank 0 has
    none: the stream starts from a reset filter), then the slice in calls of 2^29 samples with carried history.  Rank g's
    outputs are the stream's outputs [start / 4, (start + length) / 4): no collective, no exchange."""
    M, nt = w["M"], w["ntaps"]
    sl = time_slices(STREAM_SAMPLES, world, [nt], [M])[rank]
    x = torch.empty((1, sl.warmup + sl.length, 2), dtype=torch.int16, device="cuda")
    y = torch.empty((1, sl.length // M, 2), dtype=torch.int16, device="cuda")
    scratch = torch.empty((1, max(1, sl.warmup // M), 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, 0x5EED0005, ch0=0, n0=sl.start - sl.warmup, amp_shift=2)
    dec = S.FilterDnsamplingFir(M, S.design_lowpass_taps(nt, M), channels=1, device=local_rank, obsolete=True)
    dec.set_kernel(args.kernel)
    calls = [(o, min(STREAM_CALL, sl.length - o)) for o in range(0, sl.length, STREAM_CALL)]

    def step():
        dec.reset()
        if sl.warmup:
            dec.step(x[:, :sl.warmup], out=scratch)
        for o, ln in calls:
            dec.step(x[:, sl.warmup + o: sl.warmup + o + ln], out=y[:, o // M:(o + ln) // M])

    m = timed(step, args.steps, args.warmup)
    n_out_total = STREAM_SAMPLES // M
    roof = roofline_of("cfg5", w, dec, sl.length // M, m["k_ms"], m["clocks"])
    roof["launches_per_step"] = len(calls) + (1 if sl.warmup else 0)
    line = dict(base, scaling="strong", value=n_out_total * args.steps / (m["total_ms"] * 1e-3) / 1e6,
                ms_per_step=m["total_ms"] / args.steps, roofline=roof, cpu_baseline=None, e2e=None, clocks=m["clocks"],
                gpu_launches=m["launches"], impl="ours")
    line["config"] = {"workload": "cfg5: ONE stream of 2^32 complex samples, decimate-by-4 1023-tap FIR, time-sliced with halo",
                      "stream_samples": STREAM_SAMPLES, "slices": world, "slice_samples": sl.length, "halo_warmup_samples": sl.warmup,
                      "samples_per_call": STREAM_CALL, "ratio": M, "taps": nt, "nco_mix": False,
                      "sharding": "time slices at multiples of the decimation, warm-up halo, no collective (strong scaling)",
                      "l2": "2 GiB per call >> 126 MB L2 (no flush needed)"}
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


def group_bench(args, w, base, S, torch):
    """cfg3 / cfg5 through the single-process multi-device driver (srcdsp_group_*): pinned host buffers in and out, one
    host thread + three streams per device, outputs written to their final place -- north_star's "results are gathered to
    the host with async copies from pinned memory".  End to end by construction (value == e2e)."""
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--group drives all devices from ONE process: run it without torchrun")
    G = args.gpus
    if args.workload == "cfg5":
        C, n, M1, M2, nt, mode = 1, STREAM_SAMPLES, w["M"], 0, w["ntaps"], "slices"
        grp = S.DdcGroup(mode, list(range(G)), C, M1, S.design_lowpass_taps(nt, M1))
    elif args.workload == "cfg3":
        C, n, M1, M2, nt, mode = w["channels"], w["n"], w["M"], w["M2"], w["ntaps"], "channels"
        grp = S.DdcGroup(mode, list(range(G)), C, M1, S.design_lowpass_taps(nt, M1), M2, S.design_lowpass_taps(w["ntaps2"], M2), n_table=4096)
        grp.setFrequency((-1 + 2 * (np.arange(C) + 0.5) / C).astype(np.float32))
    else:
        raise SystemExit("--group is for the workloads that shard one job: cfg3 (channel batches), cfg5 (time slices)")
    Mt = M1 * (M2 or 1)
    hin, hout = S.PinnedBuffer(C, n), S.PinnedBuffer(C, n // Mt)
    # synthetic input, generated on device 0 and copied to the pinned buffer piece by piece
    piece = 1 << 28
    tmp = torch.empty((1, min(piece, n), 2), dtype=torch.int16, device="cuda:0")
    with torch.cuda.device(0):
        for c in range(C):
            for o in range(0, n, piece):
                ln = min(piece, n - o)
                S.synth_fill(tmp[:, :ln], SEED, ch0=c, n0=o, amp_shift=2)
                torch.cuda.synchronize()
                S._capi.check(S.lib().srcdsp_memcpy(0, hin.array[c, o:o + ln].ctypes.data, tmp.data_ptr(), ln * 4))
    del tmp
    grp.step(hin.array, out=hout.array)  # allocates the staging buffers
    grp.reset()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = S.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        grp.step(hin.array, out=hout.array)
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    lay, used = grp.layout()
    n_out = C * (n // Mt)
    value = n_out * args.e2e_steps / dt / 1e6
    line = dict(base, scaling="strong", value=value, ms_per_step=dt / args.e2e_steps * 1e3, steps=args.e2e_steps, roofline=None,
                cpu_baseline=None, clocks=clocks, gpu_launches=int(S.launch_count() - l0), impl="ours",
                e2e={"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": C * n * 4, "d2h_bytes_per_step": n_out * 4,
                     "h2d_gbs_total": C * n * 4 * args.e2e_steps / dt / 1e9,
                     "api": f"DdcGroup('{mode}').step(pinned host in, pinned host out) -> srcdsp_group_step"})
    line["config"] = {"workload": w["desc"].split(":")[0] + f" through one process driving {G} device(s)", "mode": mode, "channels": C,
                      "samples_per_channel": n, "ratio": Mt, "members": [{"device": d, "ch0": c0, "channels": nc} for d, c0, nc in lay],
                      "members_used": used, "sharding": "single process, one host thread + 3 streams per device, no collective"}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ddc", action="store_true", help="skip the fused mixer + decimator side measurement")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 IMAD kernel, 2 tcgen05 kernel")
    ap.add_argument("--group", action="store_true",
                    help="cfg3 / cfg5 through ONE process driving --gpus devices (srcdsp_group_*: per-device host threads and "
                         "streams, pinned host buffers in and out); not under torchrun")
    ap.add_argument("--h2d-ceiling", action="store_true",
                    help="with e2e: also time bare per-GPU cudaMemcpyAsync copies from pinned memory (all ranks at once), the box's ceiling for e2e")
    ap.add_argument("--taps", default="design", choices=["design", "impulse", "random"],
                    help="power experiments only: an impulse (all other taps zero) or dense random 2-digit taps instead of the designed low-pass")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    base = {"metric": "output Msamples/s", "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": w["desc"], "channels_per_gpu": w["channels"], "samples_per_channel": w["n"],
                       "ratio": w["M"], "taps": w["ntaps"], "nco_mix": w["mix"],
                       "sharding": "contiguous channel batches per rank, no collective",
                       "l2": "16 GiB batch per step >> 126 MB L2 (no flush needed)"}}

    # ------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        res, t = cpu_reference_run(w, args.steps, args.warmup, cores)
        line = dict(base, impl="reference", value=res["value"], ms_per_step=t * 1e3, cpu_baseline=res,
                    e2e={"value": res["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    gpu_launches=0)
        line["note"] = "CPU reference arm: each step is the bounded sample described in cpu_baseline.sample"
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------
    import torch
    import srcdsp_b200 as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: srcdsp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    from srcdsp_b200.sharding import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None  # pinned staging next to the GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    C, n, M, nt = w["channels"], w["n"], w["M"], w["ntaps"]
    if w["kind"] == "fifo":
        return fifo_stream_bench(args, w, base, S, torch, local_rank)
    if w["kind"] == "corr":
        return corr_bench(args, w, base, S, torch, local_rank)
    if w["kind"] == "decf":
        return decf_bench(args, w, base, S, torch, local_rank)
    if args.group:
        return group_bench(args, w, base, S, torch)
    from srcdsp_b200.sharding import channel_shard, time_slices

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step_fn, steps, warmup, settle_s=0.0):
        """W untimed + K timed calls of step_fn, CUDA events on the launching stream, barrier + synchronize on both
        sides, clocks / throttle reasons / power sampled during the timed region (rank 0), max over ranks.
        settle_s: idle time in front, so that a side measurement starts from the same power state as the headline
        does after its set-up (the boxes run into the 1 kW software power cap within ~100 ms of these kernels)."""
        if settle_s:
            barrier()
            time.sleep(settle_s)
        for _ in range(warmup):
            step_fn()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        l0 = S.launch_count()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            step_fn()
            evs[i + 1].record()
        barrier()
        launches = S.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
        total_ms = all_max(evs[0].elapsed_time(evs[-1]))
        step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        return {"total_ms": total_ms, "ms": total_ms / steps, "k_ms": float(np.mean(step_ms)), "clocks": clocks, "launches": int(launches)}

    def lo_freqs(ch: range, total: int):
        return (-1 + 2 * (np.arange(ch.start, ch.stop) + 0.5) / total).astype(np.float32)

    def build_chain(wk, ch: range, total_ch: int):
        """The banks of workload wk for the channels `ch` of `total_ch` on this rank's device."""
        Cc = len(ch)
        Mw, ntw = wk["M"], wk["ntaps"]
        if wk["kind"] == "mix":
            chain = S.Mixer(channels=Cc, device=local_rank)
            chain.setFrequency(lo_freqs(ch, total_ch))
            return chain, None, [chain]
        if wk["kind"] == "up":
            chain = S.FilterUpsamplingFir(Mw, S.design_interp_taps(ntw, Mw), channels=Cc, device=local_rank)
            return chain, None, [chain]
        taps = S.design_lowpass_taps(ntw, Mw)
        if args.taps == "impulse":
            taps = np.zeros(ntw, np.int32)
            taps[ntw // 2] = 32767
        elif args.taps == "random":
            taps = np.random.default_rng(1).integers(-32767, 32768, ntw).astype(np.int32)
        d = S.FilterDnsamplingFir(Mw, taps, channels=Cc, device=local_rank, obsolete=True)
        d.set_kernel(args.kernel)
        chain, stateful = d, [d]
        if wk["mix"]:
            mix = S.Mixer(channels=Cc, device=local_rank)
            mix.setFrequency(lo_freqs(ch, total_ch))
            d2 = None
            if wk["kind"] == "ddc2":
                d2 = S.FilterDnsamplingFir(wk["M2"], S.design_lowpass_taps(wk["ntaps2"], wk["M2"]), channels=Cc,
                                           device=local_rank, obsolete=True)
                stateful.append(d2)
            chain = S.Ddc(mix, d, d2)
        return chain, d, stateful

    def kernel_name(wk, d):
        if d is not None:
            return d.last_kernel
        return "mixer_seq_kernel" if wk["kind"] == "mix" else "up_fir4_kernel<8,true>" if wk["kind"] == "up" else "up_fir_kernel"

    peak, peak_src = peaks()

    def roofline_of(name, wk, d, n_out_rank, k_ms, clocks):
        alg_bytes = bytes_per_out(wk) * n_out_rank
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(name)
        roof = {"bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "kernel": kernel_name(wk, d),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                "binding": BINDING.get(name),
                "imad_note": "imad_frac = the same work counted as 2*taps INT32 multiply-adds per output against "
                             "148 SM x 64 IMAD/clk x sm_max_mhz: the ceiling of any CUDA-core kernel (SURVEY.md 8(d)); "
                             "the tcgen05 int8 kernel is not bound by it"}
        if wk["kind"] == "ddc2":
            roof["traffic_model_bytes_per_launch"] = traffic_model_bytes_per_out(wk) * n_out_rank
            roof["note"] = ("algorithmic bytes follow SURVEY.md 8(d) (132 B per output); the chain also writes and re-reads the "
                            "stage-1 stream (8 * M2 = 32 B per output more, traffic_model_bytes_per_launch)")
        roof["peak_note"] = ("peak = MEASURED_PEAKS.json's copy bandwidth (half reads, half writes); a decimator's traffic is "
                             "M reads per write, and a read-only TMA stream of this access pattern reaches ~7.7 TB/s on these boxes, "
                             "so frac can exceed 1; frac_of_spec is against the nominal 8 TB/s")
        # north_star quotes the roofline "against B200's ~8 TB/s": the same achieved rate against the nominal figure as well
        roof["spec_peak"] = 8000.0
        roof["frac_of_spec"] = roof["achieved"] / 8000.0
        if clocks and clocks.get("sm_max_mhz"):
            imad_peak = 148 * 64 * clocks["sm_max_mhz"] * 1e6
            ntw, Mw = wk["ntaps"], wk["M"]
            macs = {"dec": 2 * ntw, "ddc": 2 * ntw + 4 * Mw, "up": 2 * ntw / Mw, "mix": 4,
                    "ddc2": wk.get("M2", 1) * (2 * ntw + 4 * Mw) + 2 * wk.get("ntaps2", 0)}[wk["kind"]]
            roof["imad_frac"] = macs * n_out_rank / (k_ms * 1e-3) / imad_peak
        return roof

    # ------------------------------------------------------------------------------------------
    # multi-GPU workloads that SHARD a fixed job (strong scaling): cfg3 = 1024 channels dealt to the ranks in
    # contiguous batches; cfg5 = ONE 2^32-sample stream cut into time slices with a warm-up halo
    # ------------------------------------------------------------------------------------------
    if args.workload == "cfg5":
        return stream_bench(args, w, base, S, torch, dist, rank, world, local_rank, timed, roofline_of, time_slices)
    strong = args.workload == "cfg3"
    total_ch = C if strong else C * world
    my_ch = channel_shard(total_ch, world, rank)
    Cr = len(my_ch)
    n_out = int(n * out_per_in(w))
    x = torch.empty((Cr, n, 2), dtype=torch.int16, device="cuda")
    y = torch.empty((Cr, n_out, 2), dtype=torch.int16, device="cuda")
    S.synth_fill(x, SEED, ch0=my_ch.start, amp_shift=2)
    chain, dec, stateful = build_chain(w, my_ch, total_ch)
    m = timed(lambda: chain.step(x, out=y), args.steps, args.warmup)
    n_out_total = total_ch * n_out
    value = n_out_total * args.steps / (m["total_ms"] * 1e-3) / 1e6
    roof = roofline_of(args.workload, w, dec, Cr * n_out, m["k_ms"], m["clocks"])
    base["scaling"] = "strong" if strong else "weak"
    if strong:
        base["config"]["channels_total"] = total_ch
        base["config"]["channels_per_gpu"] = Cr
        base["config"]["sharding"] = "1024 channels in contiguous batches per rank (strong scaling), no collective"

    # ---- side measurements on the same device-resident batch (kernel-only, each with its own clocks record):
    #      ddc16 = north_star's "256-channel mix + decimate-by-16 DDC"; cfg3 / cfg4 / cfg5 = the other BASELINE configs
    #      at their single-GPU shapes; "sustained" = the headline and ddc16 over a run long enough for the power cap ----
    side, ddc_extra = {}, None
    if args.workload == "cfg2" and not args.no_ddc:
        def side_run(name, views, steps=None, warmup=None, settle=1.0):
            wk = WORKLOADS[name]
            ch = channel_shard(wk["channels"] * world, world, rank)
            try:
                xs, ys = views(wk)
                c2, d2, _ = build_chain(wk, ch, wk["channels"] * world)
                mm = timed(lambda: c2.step(xs, out=ys), steps or args.steps, warmup or args.warmup, settle)
                nout = wk["channels"] * int(wk["n"] * out_per_in(wk))
                rf = roofline_of(name, wk, d2, nout, mm["k_ms"], mm["clocks"])
                return {"workload": wk["desc"], "value": nout * world / (mm["ms"] * 1e-3) / 1e6, "unit": "Msamples/s",
                        "ms_per_step": mm["ms"], "steps": steps or args.steps, "kernel": rf["kernel"], "roofline_frac": rf["frac"],
                        "frac_of_spec": rf["frac_of_spec"], "binding": rf["binding"], "clocks": mm["clocks"],
                        "algorithmic_bytes_per_launch": rf["algorithmic_bytes_per_launch"],
                        "settle_s": settle}
            except Exception as ex:  # report, never fake
                return {"value": None, "error": str(ex)[:200]}

        flat = x.view(-1)  # the 16 GiB batch, re-viewed for the other shapes (fresh synthetic data where the shape differs)

        def in_view(wk):
            return flat[: wk["channels"] * wk["n"] * 2].view(wk["channels"], wk["n"], 2)

        def dec_views(wk):
            xs = in_view(wk)
            if (wk["channels"], wk["n"]) != (C, n):
                S.synth_fill(xs, SEED, ch0=0, amp_shift=2)
            return xs, y.view(-1)[: wk["channels"] * int(wk["n"] * out_per_in(wk)) * 2].view(wk["channels"], -1, 2)

        def up_views(wk):
            xs = in_view(wk)
            return xs, torch.empty((wk["channels"], wk["n"] * wk["M"], 2), dtype=torch.int16, device="cuda")

        ddc_extra = side_run("ddc16", dec_views)
        ddc_extra["note"] = ("same input batch and algorithmic bytes as the headline; the mixer adds 4 int32 multiply-adds, a shift "
                             "and a saturating pack per input sample in the kernel's load stage")
        side["cfg3"] = side_run("cfg3", dec_views)
        side["cfg5_slice"] = side_run("cfg5", dec_views)
        side["cfg4"] = side_run("cfg4", up_views)
        S.synth_fill(x, SEED, ch0=my_ch.start, amp_shift=2)  # the headline's batch again (for e2e)
        sus = max(100, 4 * args.steps)
        side["sustained"] = {
            "note": f"{sus} steps back to back (~0.4 s): long enough for the boxes' 1 kW software power cap to lower the SM clock; "
                    "the numbers above are the first tens of milliseconds after an idle second",
            "cfg2": side_run("cfg2", dec_views, steps=sus, settle=0.0),
            "ddc16": side_run("ddc16", dec_views, steps=sus, settle=0.0)}

    # ---- e2e: public API with pinned HOST buffers, H2D + kernels + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        try:
            # every rank pins its whole batch (17 GiB for cfg2); with many ranks on one host check that it fits
            try:
                import psutil
                need = world * (Cr * n + Cr * n_out) * 4
                if psutil.virtual_memory().available < 1.3 * need:
                    raise MemoryError(f"host has {psutil.virtual_memory().available >> 30} GiB available, "
                                      f"{need >> 30} GiB of pinned staging needed for {world} ranks")
            except ImportError:
                pass
            hin = S.PinnedBuffer(Cr, n)
            hout = S.PinnedBuffer(Cr, n_out)
            S._capi.check(S.lib().srcdsp_memcpy(local_rank, hin.array.ctypes.data, x.data_ptr(), Cr * n * 4))
            del x
            torch.cuda.empty_cache()
            chain.step(hin.array, out=hout.array)  # warm the staging buffers
            for f in stateful:
                f.reset()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                chain.step(hin.array, out=hout.array)  # returns when the output is in host memory
            barrier()
            dt = all_max(time.perf_counter() - t0)
            e2e = {"value": n_out_total * args.e2e_steps / dt / 1e6, "unit": "Msamples/s",
                   "h2d_bytes_per_step": Cr * n * 4, "d2h_bytes_per_step": Cr * n_out * 4,
                   "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
                   "h2d_gbs_per_gpu": Cr * n * 4 * args.e2e_steps / dt / 1e9,
                   "api": type(chain).__name__ + ".step(host numpy view of pinned memory) -> C ABI step"}
            if numa:
                e2e["numa_binding_rank0"] = numa
            if args.h2d_ceiling:
                e2e["h2d_ceiling"] = h2d_ceiling(torch, S, local_rank, dist, barrier, all_max)
            hin.free()
            hout.free()
        except Exception as ex:  # report, never fake
            e2e = {"value": None, "unit": "Msamples/s", "error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu, _ = cpu_reference_run(w, 2, 1, cores)
        except Exception as ex:
            cpu = {"value": None, "error": str(ex)[:200]}

    if rank == 0:
        line = dict(base, value=value, ms_per_step=m["total_ms"] / args.steps, roofline=roof, cpu_baseline=cpu, e2e=e2e,
                    clocks=m["clocks"], gpu_launches=m["launches"], impl="ours")
        if ddc_extra is not None:
            line["ddc16"] = ddc_extra
            line["side"] = side
        if args.workload == "cfg1" and world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline_variants"] = cpu_cfg1_variants(w)
            except Exception as ex:
                line["cpu_baseline_variants"] = {"error": str(ex)[:200]}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
