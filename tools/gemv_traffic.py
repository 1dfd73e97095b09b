"""profiles/gemv_traffic.json from an `ncu --set full` capture of the decode GEMV launches of bench.py (run here, no GPU needed).

usage: python tools/gemv_traffic.py gpurun_out/prof_gemv.ncu-rep llama-2-7b/world1 [commit]
Writes the average of dram__bytes_read.sum + dram__bytes_write.sum per launch over the captured w8a16_gemv_kernel launches (the four
fused decode GEMVs of whole layers, so the average is the per-launch figure bench.py's roofline uses)."""
import csv
import io
import json
import os
import subprocess
import sys

rep, key = sys.argv[1], sys.argv[2]
commit = sys.argv[3] if len(sys.argv) > 3 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n, per = 0.0, 0, []
for r in data:
    if "w8a16_gemv_kernel" not in r[col["Kernel Name"]]:
        continue
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[col[m]].replace(",", "")) * scale[units[col[m]]]
    per.append(b)
    tot += b
    n += 1
assert n > 0 and n % 4 == 0, f"expected whole layers (4 GEMV launches each), got {n} launches"
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "gemv_traffic.json")
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
out[key] = {"bytes_per_launch": tot / n, "launches": n, "per_launch": per, "source": f"ncu --set full ({os.path.basename(rep)})", "commit": commit}
json.dump(out, open(out_path, "w"), indent=1)
print(key, out[key]["bytes_per_launch"], "bytes/launch over", n, "launches")
