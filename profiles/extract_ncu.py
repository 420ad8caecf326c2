#!/usr/bin/env python
"""profiles/extract_ncu.py -- turn `ncu -i X.ncu-rep --page raw --csv` dumps of `--set full` captures into
(a) a compact per-launch summary CSV (the columns the DESIGN/README readings quote) and
(b) profiles/ncu_kernel_metrics.json, which bench.py reads for the roofline's `traffic` / FP64-pipe fields.

    python profiles/extract_ncu.py --tag r2 --batch 10000000 --commit <sha> raw1.csv [raw2.csv ...]

A kernel's entry is keyed by its bare name (template arguments and signature stripped); when several launches of one
kernel are in the dumps the one with the longest duration (the full-size launch) is kept.
"""
import argparse
import csv
import json
import os
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "duration_us"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_limit_regs_blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch_resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
              "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def bare(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"^(ncb::)", "", name)
    return re.split(r"[<(]", name, maxsplit=1)[0].replace("ncb::", "")


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:  # noqa: BLE001
        return None


def read_raw(path):
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    # skip any ==PROF== preamble
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"kernel_full": d.get("Kernel Name", ""), "kernel": bare(d.get("Kernel Name", "")),
               "block": d.get("Block Size"), "grid": d.get("Grid Size"), "source": os.path.basename(path)}
        for col, key in COLS:
            v = num(d.get(col, ""))
            if v is not None and key in ("dram_read", "dram_write", "duration_us"):
                v *= UNIT_SCALE.get(u.get(col, ""), 1.0)
            rec[key] = v
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw", nargs="+")
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--batch", type=int, default=10_000_000, help="neutrons per launch sequence of the captured command")
    ap.add_argument("--commit", default="")
    ap.add_argument("--summary", default=None)
    ap.add_argument("--json", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_kernel_metrics.json"))
    ap.add_argument("--merge", action="store_true", help="keep entries already in the json for kernels not in these dumps")
    a = ap.parse_args()
    recs = []
    for p in a.raw:
        recs += read_raw(p)
    if not recs:
        sys.exit("no launches found")
    keys = ["kernel_full", "block", "grid"] + [k for _, k in COLS] + ["source"]
    summ = a.summary or os.path.join(os.path.dirname(a.json), "%s_kernels_ncu_full.csv" % a.tag)
    with open(summ, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keys)
        for r in recs:
            w.writerow([r.get(k) for k in keys])
    best = {}
    for r in recs:
        if r["duration_us"] is None:
            continue
        if r["kernel"] not in best or r["duration_us"] > best[r["kernel"]]["duration_us"]:
            best[r["kernel"]] = r
    js = {}
    if a.merge and os.path.exists(a.json):
        js = json.load(open(a.json))
    for k, r in best.items():
        dr, dw = r.get("dram_read") or 0.0, r.get("dram_write") or 0.0
        js[k] = {"capture": "profiles/" + os.path.basename(summ), "commit": a.commit, "batch_neutrons": a.batch,
                 "kernel_full": r["kernel_full"], "duration_us_under_ncu": r["duration_us"],
                 "dram_bytes_per_launch": dr + dw, "dram_read_bytes": dr, "dram_write_bytes": dw,
                 "fp64_pipe_pct": r["fp64_pipe_pct"], "issue_active_pct": r["issue_active_pct"], "l2_hit_pct": r["l2_hit_pct"],
                 "l1_hit_pct": r["l1_hit_pct"], "lanes_per_inst": r["lanes_per_inst"],
                 "achieved_occupancy_pct": r["achieved_occupancy_pct"], "regs": r["regs"]}
    json.dump(js, open(a.json, "w"), indent=1, sort_keys=True)
    print("wrote", summ, "and", a.json, "(%d kernels)" % len(best))


if __name__ == "__main__":
    main()
