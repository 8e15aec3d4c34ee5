"""profiles/roofline_latest.json from an ncu launch list that carries gpu__time_duration.sum, dram__bytes_read.sum and
dram__bytes_write.sum per launch (`ncu --metrics ... --csv --log-file x.csv <command>`):

    python tools/make_roofline.py profiles/launches_r01_final.csv "cornell 1024x1024 1 pass"

bench.py reads `shade.dram_bytes_per_launch` / `extend.dram_bytes_per_launch` as roofline.traffic."""
import csv
import json
import sys
from collections import defaultdict
from pathlib import Path

src = Path(sys.argv[1]); what = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(l for l in src.read_text().splitlines() if l.startswith('"')))
h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = defaultdict(dict)
for r in rows[1:]:
    per[(int(r[ii]), r[ki])][r[mi]] = float(r[vi].replace(",", ""))
kern = defaultdict(lambda: dict(launches=0, time_ms=0.0, dram=0.0))
for (_, name), m in per.items():
    short = name.split("(")[0].replace("void ", "").replace("iptd::", "")
    k = kern[short]
    k["launches"] += 1; k["time_ms"] += m.get("gpu__time_duration.sum", 0) / 1e6
    k["dram"] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
total = sum(k["time_ms"] for k in kern.values())
out = {"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, {what} ({src})", "kernels": {}}
for name, k in sorted(kern.items(), key=lambda kv: -kv[1]["time_ms"]):
    out["kernels"][name] = {"launches": k["launches"], "time_ms": k["time_ms"], "share": k["time_ms"] / total,
                            "dram_bytes_per_launch": k["dram"] / k["launches"]}
for group, prefix in (("shade", "k_shade"), ("extend", "k_extend")):
    sel = [k for n, k in kern.items() if n.startswith(prefix)]
    if sel:
        n = sum(k["launches"] for k in sel)
        out[group] = {"launches": n, "time_ms": sum(k["time_ms"] for k in sel), "share": sum(k["time_ms"] for k in sel) / total,
                      "dram_bytes_per_launch": sum(k["dram"] for k in sel) / n}
Path("profiles/roofline_latest.json").write_text(json.dumps(out, indent=1))
print(json.dumps({g: out[g] for g in ("shade", "extend") if g in out}))
