"""Per-region instruction and stall-sample totals of an ncu capture (SASS view): instructions are bucketed by how often
they executed, which separates the per-step loops from the per-chunk passes.  usage: ncu_regions.py report.ncu-rep warp_steps"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]; steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if r and r[0] == 'Address': hdr = r; start = i; break
ii = hdr.index('Instructions Executed'); wi = hdr.index('Warp Stall Sampling (All Samples)'); si = hdr.index('Source')
data = [(r[si], int(r[ii] or 0), int(r[wi] or 0)) for r in rows[start + 1:] if len(r) > wi]
tot = sum(d[2] for d in data); ti = sum(d[1] for d in data)
print("warp-instructions %.1fM = %.1f per warp-step; stall samples %d" % (ti / 1e6, ti / steps, tot))
b = Counter(); s = Counter(); n = Counter()
for d in data:
    e = d[1]
    key = 0 if e == 0 else round(e, -len(str(e)) + 2)
    b[key] += e; s[key] += d[2]; n[key] += 1
for key in sorted(b, key=lambda k: -b[k])[:14]:
    print("exec ~%9d: %4d instrs, %6.1fM executed (%5.1f per warp-step), %4.1f%% of stall samples" % (key, n[key], b[key] / 1e6, b[key] / steps, 100 * s[key] / max(tot, 1)))
