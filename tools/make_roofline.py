"""profiles/roofline_latest.json from ncu output of `tools/profile_run.py <workload>`:

    python tools/make_roofline.py <workload>=<launch list csv>[,<raw page csv of a --set full capture>] ...

launch list: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
--log-file x.csv python tools/profile_run.py c2 2` (per-launch times are cold-cache and serialised: the SHARE of the step
is what must agree with bench.py's CUDA-event times). raw page: `ncu -i x.ncu-rep --page raw --csv` of a `--set full`
capture of the dominant launches (issue-slot utilisation, active lanes per instruction, L2 hit rate).

bench.py reads workloads[<workload>][shade|extend]: dram_bytes_per_launch -> roofline.traffic, issue_active_pct ->
roofline.issue_frac, lanes_per_inst. The first launch of every kernel group is the warm-up pass of profile_run.py and is
left out when a second one exists."""
import csv
import hashlib
import json
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def short(name):
    return name.split("(")[0].replace("void ", "").replace("iptd::", "")


def launch_list(path):
    rows = list(csv.reader(l for l in Path(path).read_text().splitlines() if l.startswith('"')))
    h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    ui = h.index("Metric Unit")
    per = defaultdict(dict)
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)  # -> ms
        elif r[mi].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        per[(int(r[ii]), short(r[ki]))][r[mi]] = v
    return [(i, n, m) for (i, n), m in sorted(per.items())]


def raw_page(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    out = defaultdict(list)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        def f(k):
            try: return float(d[k].replace(",", ""))
            except Exception: return None
        out[short(d["Kernel Name"])].append(dict(
            issue_active_pct=f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            lanes_per_inst=f("smsp__thread_inst_executed_per_inst_executed.ratio"),
            l2_hit_pct=f("lts__t_sector_hit_rate.pct"), l1_hit_pct=f("l1tex__t_sector_hit_rate.pct"),
            registers=f("launch__registers_per_thread"), warps_active_pct=f("sm__warps_active.avg.pct_of_peak_sustained_active"),
            warp_instructions=f("smsp__inst_executed.sum"), time_ms_under_full_set=f("gpu__time_duration.sum"),
            long_scoreboard_per_issue=f("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            l1tex_throughput_pct=f("l1tex__throughput.avg.pct_of_peak_sustained_active")))
    return out


def main():
    h = hashlib.sha256()
    for p in sorted((ROOT / "ipt_b200" / "csrc").glob("*.cu*")):
        h.update(p.read_bytes())
    out = {"source": "ncu launch lists (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none) and --set full "
                     "raw pages of `python tools/profile_run.py <workload>`; profile-time figures, never a bench value", "source_sha16": h.hexdigest()[:16], "workloads": {}}
    for arg in sys.argv[1:]:
        wl, files = arg.split("=")
        files = files.split(",")
        launches = launch_list(files[0])
        raw = raw_page(files[1]) if len(files) > 1 else {}
        groups = defaultdict(list)
        for i, n, m in launches:
            groups[n].append(m)
        kern = {}
        for n, ms in groups.items():
            use = ms[len(ms) // 2:] if len(ms) > 1 else ms  # second half = the measured pass (first half = warm-up pass)
            kern[n] = dict(launches=len(use), time_ms=sum(m.get("gpu__time_duration.sum", 0) for m in use),
                           dram_bytes=sum(m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0) for m in use))
        total = sum(k["time_ms"] for k in kern.values())
        rec = {"files": files, "kernels": {}}
        for n, k in sorted(kern.items(), key=lambda kv: -kv[1]["time_ms"]):
            e = dict(launches=k["launches"], time_ms=k["time_ms"], share=k["time_ms"] / total, dram_bytes_per_launch=k["dram_bytes"] / k["launches"])
            if n in raw:
                e["full_set"] = raw[n]
            rec["kernels"][n] = e
        for group, prefix in (("shade", "k_shade"), ("extend", "k_extend")):
            sel = {n: k for n, k in kern.items() if n.startswith(prefix)}
            if not sel:
                continue
            nl = sum(k["launches"] for k in sel.values())
            g = dict(launches=nl, time_ms=sum(k["time_ms"] for k in sel.values()), share=sum(k["time_ms"] for k in sel.values()) / total,
                     dram_bytes_per_launch=sum(k["dram_bytes"] for k in sel.values()) / nl)
            caps = [c for n in sel for c in raw.get(n, []) if c.get("issue_active_pct") is not None]
            if caps:  # time-weighted over the captured launches of the group
                wsum = sum(c["time_ms_under_full_set"] or 1.0 for c in caps)
                g["issue_active_pct"] = sum(c["issue_active_pct"] * (c["time_ms_under_full_set"] or 1.0) for c in caps) / wsum
                g["lanes_per_inst"] = sum(c["lanes_per_inst"] * (c["time_ms_under_full_set"] or 1.0) for c in caps) / wsum
                g["l2_hit_pct"] = sum((c["l2_hit_pct"] or 0) * (c["time_ms_under_full_set"] or 1.0) for c in caps) / wsum
            rec[group] = g
        out["workloads"][wl] = rec
    (ROOT / "profiles" / "roofline_latest.json").write_text(json.dumps(out, indent=1))
    print(json.dumps({wl: {g: r[g] for g in ("shade", "extend") if g in r} for wl, r in out["workloads"].items()}, indent=1))


if __name__ == "__main__":
    main()
