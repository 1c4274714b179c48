"""Summarise one kernel launch of an .ncu-rep (raw page) into the JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/r1_batch8.ncu-rep profiles/r1_icp_kernel_batch8_summary.json "note ..."
"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__warps_active.avg.per_cycle_active",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": vals[i], "unit": units[i]}
    by = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        by += float(d[k]["value"]) * UNIT[d[k]["unit"]]
    d["dram_bytes_per_launch"] = by
    d["kernel"] = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    d["note"] = note
    json.dump(d, open(out, "w"), indent=1)
    print(out, "dram bytes/launch", by, "time", d["gpu__time_duration.sum"])


if __name__ == "__main__":
    main()
