#!/usr/bin/env python
"""Reads an .ncu-rep (one `ncu --set full` capture of one launch on the benchmark batch) and

  * writes the selected metrics as JSON next to the other round summaries (profiles/<name>.json), and
  * updates profiles/current.json, the record bench.py quotes in its roofline object:
        { workload: { kernel: { "dram_bytes": read + write per launch, "inst_executed": warp instructions per launch,
                                "duration_ms": ..., "file": "profiles/<name>.json" } } }

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r2_ncu_full_lockstep_1024x1080p.json yuvf vp8_mb_lockstep
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "sm__cycles_elapsed.max", "sm__inst_executed.sum",
        "smsp__inst_executed.sum", "sm__inst_issued.avg.per_cycle_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_bytes.sum", "sm__icc_requests_lookup_hit.sum",
        "sm__icc_requests_lookup_miss.sum", "launch__shared_mem_per_block_dynamic")


def raw_metrics(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", f"regex:{kernel}"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]  # first matching launch
    return {h: {"unit": u, "value": v} for h, u, v in zip(hdr, units, vals)}


def num(m, key):
    v = float(m[key]["value"].replace(",", ""))
    unit = m[key]["unit"].lower()
    scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit, 1)
    return v * scale


if __name__ == "__main__":
    rep, dst, workload, kernel = sys.argv[1:5]
    m = raw_metrics(rep, kernel)
    keep = {k: v for k, v in m.items() if k in KEEP or k.startswith("smsp__average_warp") or k.startswith("smsp__average_warps_issue_stalled")}
    keep["Kernel Name"] = m.get("Kernel Name", {"value": kernel})
    Path(dst).write_text(json.dumps(keep, indent=1, sort_keys=True))
    cur_p = ROOT / "profiles" / "current.json"
    cur = json.loads(cur_p.read_text()) if cur_p.exists() else {}
    inst_key = "smsp__inst_executed.sum" if "smsp__inst_executed.sum" in m else "sm__inst_executed.sum"
    cur.setdefault(workload, {})[kernel] = {"dram_bytes": num(m, "dram__bytes_read.sum") + num(m, "dram__bytes_write.sum"),
                                            "dram_bytes_read": num(m, "dram__bytes_read.sum"), "dram_bytes_write": num(m, "dram__bytes_write.sum"),
                                            "inst_executed": num(m, inst_key), "duration_ms": num(m, "gpu__time_duration.sum"),
                                            "file": str(Path(dst).resolve().relative_to(ROOT))}
    cur_p.write_text(json.dumps(cur, indent=1, sort_keys=True))
    print(json.dumps(cur[workload][kernel]))
