#!/usr/bin/env python
"""Summarise one `ncu --set full` report into the three files profiles/ keeps per capture:

    <out>_ncu_metrics.txt     the metrics the judge greps for (time, pipes, issue, stalls, occupancy, DRAM bytes)
    <out>_source_digest.txt   stall reasons, executed mix, hottest instructions (tools/ncu_source.py)
    <out>_ncu.json            sidecar bench.py reads for roofline.traffic: kernel, workload, DRAM bytes per launch

Usage: ncu_digest.py report.ncu-rep profiles/r02_f32_K19W8_cfg2 --config 2 --scale 1.0 --role f32|f64 [--note "..."]
Runs in the build container (ncu reads reports without a GPU)."""
import argparse, csv, datetime, io, json, os, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("out")
ap.add_argument("--config", type=int, required=True); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--role", required=True, help="f32 | f64 | sw ...")
ap.add_argument("--note", default="")
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True).stdout.decode("utf-8", "replace")
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, val = rows[0], rows[1], rows[-1]
m = {h: (u, v) for h, u, v in zip(hdr, units, val)}
KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct"]
KEEP += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))


def to_bytes(u, v):
    x = float(v.replace(",", "")) if v not in ("", "n/a") else 0.0
    return int(x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1))


with open(a.out + "_ncu_metrics.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none; {a.note}\n")
    for k in KEEP:
        if k in m:
            f.write(f"{k:105s} {m[k][0]:16s} {m[k][1]}\n")
dur_u, dur_v = m.get("gpu__time_duration.sum", ("ms", "0"))
dur_ms = float(dur_v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(dur_u, 1.0)
side = {"kernel": m.get("Kernel Name", ("", ""))[1], "role": a.role, "config": a.config, "scale": a.scale,
        "dram_bytes_read": to_bytes(*m.get("dram__bytes_read.sum", ("byte", "0"))),
        "dram_bytes_write": to_bytes(*m.get("dram__bytes_write.sum", ("byte", "0"))),
        "duration_ms": dur_ms, "registers": int(float(m.get("launch__registers_per_thread", ("", "0"))[1] or 0)),
        "report": os.path.basename(a.report), "note": a.note,
        "captured": datetime.datetime.fromtimestamp(os.path.getmtime(a.report)).isoformat(timespec="seconds")}
with open(a.out + "_ncu.json", "w") as f:
    json.dump(side, f, indent=1)
src = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True).stdout.decode("utf-8", "replace")
tmp = a.out + "_src.tmp.csv"
open(tmp, "w").write(src)
dig = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_source.py"), tmp, "40"],
                     capture_output=True, text=True).stdout
os.remove(tmp)
with open(a.out + "_source_digest.txt", "w") as f:
    f.write(f"# {a.note}\n" + dig)
print(json.dumps(side))
