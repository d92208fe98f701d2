"""Top SASS lines by warp-stall samples of one captured kernel (ncu --set full --import-source on):
usage: ncu_stalls.py <report.ncu-rep> [top_n]   (runs `ncu -i ... --page source --csv` and sorts by samples)"""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r))
h = rows[hi]
si = h.index("Source")
sa = next(i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Samples" in c)
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_")]
recs = []
for r in rows[hi + 1:]:
    try:
        n = int(r[sa])
    except (ValueError, IndexError):
        continue
    recs.append((n, r))
tot = sum(n for n, _ in recs)
print(f"{len(recs)} SASS lines, {tot} samples; sample column: {h[sa]}")
for n, r in sorted(recs, key=lambda t: -t[0])[:top]:
    why = sorted(((int(r[i]) if r[i].isdigit() else 0, h[i]) for i in stall_cols), reverse=True)[:3]
    print(f"{n:7d} {100.0 * n / max(tot, 1):5.1f}%  {r[si][:90]:90s} " + " ".join(f"{w}:{c}" for c, w in why if c))
