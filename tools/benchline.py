"""Print the interesting fields of bench.py's JSON line (stdin)."""
import json
import sys

for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    print(f"{d.get('impl')} {d['config']['workload'][:6]} n_gpus={d['n_gpus']} value={d['value']:.1f} {d['unit']} "
          f"ms/step={d['ms_per_step']:.3f} hbm_frac={r.get('frac')} achieved={r.get('achieved')} "
          f"e2e={e.get('value')} cpu={c.get('value')} clocks={d.get('clocks')} launches={d.get('gpu_launches')}")
