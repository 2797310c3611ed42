"""profiles/*_ncu_summary.json from an ncu report of ONE launch of the step kernel.

    ncu --set full --clock-control none --import-source on -k regex:jb_sim_kernel -s 757 -c 1 -o gpurun_out/X \
        python tests/profile_step.py 1048576 150            (on the GPU box, after the same command exited 0 without ncu)
    python tests/scripts/ncu_summary.py gpurun_out/X.ncu-rep profiles/r02_step_ncu_summary.json [envs] [launches per control step]

Reads the report with `ncu -i ... --page raw --csv` (the metric list below) and `--page source --csv` (warp stall
samples by reason, opcode mix); keeps the keys `kernel`, `command`, `history` and `reading` of an existing output file.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
    "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads",
    "sass__inst_executed_global_stores",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, out = sys.argv[1], sys.argv[2]
    envs = int(sys.argv[3]) if len(sys.argv) > 3 else 1048576
    per_step = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    raw = ncu_csv(rep, "raw")
    hdr, units, row = raw[0], raw[1], raw[2]
    metrics = {}
    for h, u, v in zip(hdr, units, row):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            if num(v) is not None:
                metrics[h] = [num(v), u]
    dram = sum(metrics[k][0] * UNIT.get(metrics[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    src = ncu_csv(rep, "source")
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    stall = collections.Counter()
    ops, lanes, total = collections.Counter(), collections.Counter(), 0.0
    for r in src[2:]:
        if len(r) < len(h2):
            continue
        for k in ix:
            if k.startswith("stall_") and "(" not in k and num(r[ix[k]]):
                stall[k] += num(r[ix[k]])
        n, t = num(r[ix["Instructions Executed"]]), num(r[ix["Thread Instructions Executed"]])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        if n and m:
            ops[m.group(2)] += n
            lanes[m.group(2)] += t or 0.0
            total += n
    ssum = sum(stall.values()) or 1.0
    prev = json.load(open(out)) if os.path.exists(out) else {}
    doc = {
        "kernel": prev.get("kernel", "jb_sim_kernel"), "command": prev.get("command", ""), "envs": envs,
        "launches_per_control_step": per_step, "dram_traffic_bytes_per_launch": dram,
        "dram_traffic_bytes_per_env_per_launch": dram / envs,
        "dram_traffic_bytes_per_env_per_control_step": dram / envs * per_step,
        "algorithmic_bytes_per_env_per_control_step": 460,
        "metrics": metrics,
        "warp_stall_samples_pct": {k: round(100 * v / ssum, 2) for k, v in stall.most_common()},
        "opcode_mix_pct_of_warp_instructions (active lanes)": {k: [round(100 * v / total, 1), round(lanes[k] / v, 1)]
                                                               for k, v in ops.most_common(16)},
        "history": prev.get("history", {}), "reading": prev.get("reading", ""),
    }
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(out, "%.2f ms" % metrics["gpu__time_duration.sum"][0], "issue active %.1f %%" %
          metrics["smsp__issue_active.avg.pct_of_peak_sustained_active"][0], "DRAM %.2f GB per launch" % (dram / 1e9))


if __name__ == "__main__":
    main()
