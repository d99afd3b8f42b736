"""Reads .ncu-rep files brought back in gpurun_out/ (here, no GPU needed: `ncu -i ... --page raw --csv`) and writes what the judge
and bench.py read: profiles/ncu_<name>_raw.csv (the whole raw page) and the per-launch figures in profiles/roofline_traffic.json.

  python scripts/ncu_extract.py gpurun_out/prof_trace_r02.ncu-rep sponza_1080p trace
  python scripts/ncu_extract.py gpurun_out/prof_trace10m_r02.ncu-rep instanced_10m trace10m
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "smsp__inst_executed.sum", "lts__t_bytes.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def unit_scale(unit, name):
    u = (unit or "").lower()
    if name.startswith("gpu__time_duration"):
        return {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}.get(u, 1.0)   # -> ms
    if "byte" in u:
        return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    return 1.0


def main():
    rep, workload, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    out_csv = os.path.join(ROOT, "profiles", f"ncu_{tag}_r02_raw.csv")
    open(out_csv, "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(head)}
    summary = []
    for r in body:
        rec = {"id": r[idx["ID"]], "kernel": r[idx["Kernel Name"]][:90]}
        for k in KEEP:
            if k in idx:
                try:
                    rec[k] = float(r[idx[k]].replace(",", "")) * unit_scale(units[idx[k]], k)
                except ValueError:
                    rec[k] = r[idx[k]]
        summary.append(rec)
    for rec in summary:
        ms = rec.get("gpu__time_duration.sum", 0.0)
        dram = rec.get("dram__bytes_read.sum", 0.0) + rec.get("dram__bytes_write.sum", 0.0)
        rec["dram_bytes"] = dram
        rec["dram_gbs"] = dram / (ms * 1e-3) / 1e9 if ms else None
        print(json.dumps(rec))
    # the first captured launch is the bounce-0 closest-hit kernel of the first timed batch (see scripts/gpu_r02_b.sh)
    traffic_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        traffic = json.load(open(traffic_path))
    except Exception:
        traffic = {}
    first = summary[0]
    traffic[f"{workload}_closest_dram_bytes_per_launch"] = first["dram_bytes"]
    traffic[f"{workload}_closest_dram_gbs"] = first["dram_gbs"]
    traffic[f"{workload}_closest_ms_under_ncu"] = first.get("gpu__time_duration.sum")
    traffic[f"{workload}_closest_lanes_per_instruction"] = first.get("smsp__thread_inst_executed_per_inst_executed.ratio")
    traffic[f"{workload}_captures"] = [{k: v for k, v in rec.items() if k in ("id", "kernel", "gpu__time_duration.sum", "dram_bytes", "dram_gbs",
                                                                                "smsp__thread_inst_executed_per_inst_executed.ratio",
                                                                                "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct")} for rec in summary]
    traffic["source"] = "scripts/ncu_extract.py over gpurun_out/*.ncu-rep of scripts/gpu_r02_b.sh (ncu --set full --clock-control none)"
    json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
