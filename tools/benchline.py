"""Condenses bench.py's JSON line (read from stdin, other output ignored) into one short line."""
import json
import sys

for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        if ln and "NCCL" not in ln:
            print("  | " + ln[:200])
        continue
    try:
        d = json.loads(ln)
    except ValueError:
        print("  | " + ln[:200])
        continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    print(f"{d.get('impl')} {d['config']['workload'].split(':')[0]} n={d['n_gpus']} {d['value']:.0f} Msps "
          f"{d['ms_per_step']:.3f} ms frac={r.get('frac', 0):.3f} kernel={r.get('kernel')} e2e={e.get('value')} cpu={c.get('value')} "
          f"clk={(d.get('clocks') or {}).get('sm_mhz')} {(d.get('clocks') or {}).get('reasons')} launches={d.get('gpu_launches')}")
