"""Aggregate an ncu source page by CUDA source line from its SASS rows (robust to commas / quotes in the source text):
stall samples with the top stall reasons, instructions, shared-memory wavefronts (measured / ideal).
usage: python scripts/ncu_lines2.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not" not in n]
agg = collections.defaultdict(lambda: collections.Counter()); text = {}
cur = None
for r in rows:
    if not r: continue
    if r[0].isdigit():
        cur = int(r[0]); text.setdefault(cur, ",".join(r[1:6])[:90]); continue
    if cur is None or len(r) != len(hdr) or not r[ix["Address"]].startswith("0x"): continue
    def f(n):
        try: return float(r[ix[n]])
        except ValueError: return 0.0
    a = agg[cur]
    a["samples"] += f("Warp Stall Sampling (All Samples)"); a["inst"] += f("Instructions Executed")
    a["wf"] += f("L1 Wavefronts Shared"); a["wf_ideal"] += f("L1 Wavefronts Shared Ideal")
    for s in stalls: a[s] += f(s)
ts = sum(a["samples"] for a in agg.values()); ti = sum(a["inst"] for a in agg.values())
print("total warp-instructions %d, stall samples %d" % (ti, ts))
print("--- by stall samples")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    tp = sorted(((s, a[s]) for s in stalls if a[s] > 0), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% stall %5.1f%% inst  line %4d  %s | %s" % (100 * a["samples"] / max(ts, 1), 100 * a["inst"] / max(ti, 1), ln,
          " ".join("%s=%d" % (s[6:], v) for s, v in tp), text[ln]))
print("--- by shared-memory wavefronts (measured / ideal)")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["wf"])[:10]:
    print("  line %4d  %12d / %12d  %s" % (ln, a["wf"], a["wf_ideal"], text[ln]))
