"""Same-box sweeps: tools/sweep.py [--reps R] [--steps K] variant...   (run under gpurun from the repo root)

A variant is  name:lib:ENV=V,ENV=V:workload  -- `lib` is a file under build/ab/ (copied over the in-tree library
before the run; "-" keeps the current one), the environment entries are SRCDSP_* tuning knobs (read once per
process by the library).  Variants are run round-robin R times, so box drift hits all of them alike; one line per
run, then the per-variant minimum and median of ms_per_step.
"""
import json
import os
import shutil
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "srcdsp_b200", "lib", "libsrcdsp_b200.so")


def main():
    args = sys.argv[1:]
    reps, steps = 3, 10
    while args and args[0].startswith("--"):
        if args[0] == "--reps":
            reps = int(args[1])
        elif args[0] == "--steps":
            steps = int(args[1])
        args = args[2:]
    variants = []
    for a in args:
        name, lib, env, wl = a.split(":")
        wl, _, extra = wl.partition("+")  # workload+--flag=value+... : extra bench.py arguments
        variants.append((name, lib, dict(e.split("=") for e in env.split(",") if e), wl, [x for x in extra.split("+") if x]))
    res = {v[0]: [] for v in variants}
    for rep in range(reps):
        for name, lib, env, wl, extra in variants:
            if lib != "-":
                shutil.copyfile(os.path.join(ROOT, "build", "ab", lib), LIB)
            e = dict(os.environ)
            e.update(env)
            p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", str(steps), "--warmup", "3",
                                "--no-ddc", "--no-cpu", "--no-e2e"] + extra, capture_output=True, text=True, env=e, timeout=600)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            if p.returncode != 0 or not line:
                print(f"{name} rep{rep}: FAILED rc={p.returncode} {p.stderr[-300:]}", flush=True)
                continue
            for ln in p.stderr.splitlines():
                if "counters" in ln or "dec_tma_kernel:" in ln:
                    print("   | " + ln[:600], flush=True)
            d = json.loads(line[-1])
            res[name].append(d["ms_per_step"])
            print(f"{name} {wl} rep{rep}: {d['ms_per_step']:.3f} ms frac={d['roofline']['frac']:.3f} clk={d['clocks']['sm_mhz']}/{d['clocks'].get('sm_min_mhz')} "
                  f"P={d['clocks'].get('power_w')}/{d['clocks'].get('power_w_max')}W {d['clocks']['reasons']}", flush=True)
    print("---- summary (min / median ms) ----")
    for name, v in res.items():
        if v:
            print(f"{name}: {min(v):.3f} / {statistics.median(v):.3f}   n={len(v)}")


if __name__ == "__main__":
    main()
