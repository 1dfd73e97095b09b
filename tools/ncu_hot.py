"""Print the hottest SASS lines (warp-stall samples) of one kernel from an .ncu-rep (run here, no GPU needed).
usage: python tools/ncu_hot.py <report.ncu-rep> [launch index] [top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = txt.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
si, ai = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
out, tot = [], 0
for n, r in enumerate(rows[1:]):
    try:
        v = int(r[ai])
    except Exception:
        continue
    tot += v
    out.append((v, n, r[si].strip()[:120]))
print("total samples", tot, "instructions", len(out))
for v, n, s in sorted(out, reverse=True)[:top]:
    print(f"{v:6d} {100.0 * v / max(tot, 1):5.1f}%  #{n:4d}  {s}")
