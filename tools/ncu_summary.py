"""Summarise an ncu report (read here, no GPU needed) into profiles/<name>.md + ncu_traffic.json.

    python tools/ncu_summary.py gpurun_out/r01_rowteam.ncu-rep profiles/r01_score_rowteam [--traffic-key score_kernel]
"""
import csv, io, json, subprocess, sys, os

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = ["barrier", "dispatch_stall", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "not_selected",
          "wait", "selected", "mio_throttle", "lg_throttle", "no_instruction", "branch_resolving", "sleeping", "membar", "drain"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of `{os.path.basename(rep)}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200;",
             "read back with `ncu -i ... --page raw --csv` (tools/ncu_summary.py).  Per-launch values.", ""]
    summary = []
    for r in data:
        name = r[ix["Kernel Name"]]
        d = {"kernel": name}
        lines += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in ix:
                lines.append(f"| {k} | {r[ix[k]]} | {units[ix[k]]} |")
                d[k] = r[ix[k]]
        lines += ["", "| warp stall (per issue-active cycle) | ratio |", "|---|---|"]
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in ix:
                lines.append(f"| {s} | {r[ix[k]]} |")
                d["stall_" + s] = r[ix[k]]
        lines.append("")
        summary.append(d)
    with open(out + ".md", "w") as f:
        f.write("\n".join(lines))
    with open(out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()
