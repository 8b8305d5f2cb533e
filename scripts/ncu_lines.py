"""Aggregate an ncu source page (cuda,sass view) by CUDA source line.
usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""; data = []; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("", "Function Name"): continue
    try:
        ii = hdr.index("Instructions Executed"); wi = hdr.index("Warp Stall Sampling (All Samples)")
        data.append((int(r[ii] or 0), int(r[wi] or 0), cur_file, r[0], r[1].strip()[:100]))
    except Exception:
        pass
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("total warp-instructions %d, stall samples %d" % (ti, ts))
print("--- by instructions")
for d in sorted(data, key=lambda d: -d[0])[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * d[0] / max(ti, 1), 100 * d[1] / max(ts, 1), d[2], d[3], d[4]))
print("--- by stall samples")
for d in sorted(data, key=lambda d: -d[1])[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * d[0] / max(ti, 1), 100 * d[1] / max(ts, 1), d[2], d[3], d[4]))
