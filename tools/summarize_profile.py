"""Turn gpurun_out ncu artefacts into the committed text summaries under profiles/.
usage: summarize_profile.py <tag> <prof.ncu-rep> [launches.csv]"""
import csv, os, subprocess, sys
tag, rep = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 else None
os.makedirs("profiles", exist_ok=True)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
with open("profiles/%s_ncu_full_summary.txt" % tag, "w") as f:
    f.write("# ncu --set full --clock-control none, one block per captured launch (source: %s)\n" % os.path.basename(rep))
    for r in rows[2:]:
        f.write("-----\n")
        for w in want:
            if w in idx:
                f.write("%s = %s %s\n" % (w, r[idx[w]], units[idx[w]]))
if launches and os.path.exists(launches):
    lr = [r for r in csv.reader(open(launches)) if r]
    h = next(i for i, r in enumerate(lr) if "Kernel Name" in r)
    H = lr[h]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
    agg = {}
    for r in lr[h + 1:]:
        if len(r) <= mv: continue
        name = r[kn].split("(")[0]
        v = float(r[mv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open("profiles/%s_launches.txt" % tag, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# %-70s %6s %14s %7s\n" % ("kernel", "count", "total ns", "share"))
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %6d %14.0f %6.1f%%\n" % (name[:72], a[0], a[1], 100 * a[1] / tot))
print("written")
