"""Attributes shared-memory wavefronts to instructions from an `ncu --page source --csv --print-source sass` export
(profiles/*_ncu_source_*.csv.gz): for every LDS / STS with more than 1 M wavefronts, the wavefronts the instruction
caused against the ideal for its access pattern.  Excess = 0 means the "bank conflicts" of ncu's details page are not
address conflicts of the code but arbitration with the async proxy (TMA writes, tcgen05 operand reads)."""
import csv
import gzip
import sys

rows = list(csv.reader(gzip.open(sys.argv[1], "rt")))
hdr = next(r for r in rows if r and r[0] == "Address")
data = rows[rows.index(hdr) + 1:]
col = {h: i for i, h in enumerate(hdr)}
tot_w = tot_i = tot_e = 0
print(f"{'instruction':58s} {'executed':>10s} {'wavefronts':>12s} {'ideal':>12s} {'excess':>10s}")
for r in data:
    w = int(r[col["L1 Wavefronts Shared"]] or 0)
    if w < 1_000_000:
        continue
    i, e = int(r[col["L1 Wavefronts Shared Ideal"]] or 0), int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
    tot_w, tot_i, tot_e = tot_w + w, tot_i + i, tot_e + e
    print(f"{r[col['Source']].strip()[:58]:58s} {int(r[col['Instructions Executed']]):10d} {w:12d} {i:12d} {e:10d}")
print(f"{'TOTAL (instructions above 1 M wavefronts)':58s} {'':10s} {tot_w:12d} {tot_i:12d} {tot_e:10d}")
