"""Turn one `ncu --set full --import-source on` report into the text summary kept under profiles/.

    python profiles/ncu_summary.py gpurun_out/prof_r2_ls4.ncu-rep > profiles/r2_ls4_history_kernel_summary.txt

Prints the launch's key counters (`--page raw`), the stall mix per issued instruction, and the source lines with the most
warp-stall samples together with their share of executed warp instructions and their average active lanes
(`--page source --print-source cuda,sass`)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_config_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_average_branch_targets_threads_uniform.pct",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def main(rep, top=28):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
    print("kernel:", d.get("Kernel Name", ("?",))[0])
    for k in KEYS:
        if k in d:
            print("%-72s %s %s" % (k, d[k][0], d[k][1]))
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = float(d[h][0] or 0)
            if v >= 0.04:
                print("  stall %-28s %.3f per issued instruction" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur, agg = None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0].isdigit() and len(r) > 8:
            f = lambda x: int(x) if x.isdigit() else 0
            agg.append((cur, int(r[0]), r[1].strip()[:104], f(r[6]), f(r[7]), f(r[8])))
    ts, ti, tt = sum(a[3] for a in agg), sum(a[4] for a in agg), sum(a[5] for a in agg)
    print("warp-stall samples %d, warp instructions (source-correlated) %d, average active lanes %.2f" % (ts, ti, tt / max(ti, 1)))
    serial = [a for a in agg if a[4] and a[5] / a[4] <= 8.5]
    print("lines that run on at most 8 lanes (the one-thread-per-history phase): %.1f %% of samples, %.1f %% of warp instructions"
          % (100.0 * sum(a[3] for a in serial) / ts, 100.0 * sum(a[4] for a in serial) / ti))
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        print("%-18s %4d smp %5.2f%% inst %5.2f%% lanes %4.1f  %s" % (a[0], a[1], 100.0 * a[3] / ts, 100.0 * a[4] / ti, a[5] / max(a[4], 1), a[2]))


if __name__ == "__main__":
    main(sys.argv[1])
