"""Key metrics of every captured launch of an ncu report (ncu --set full), as the text summaries under profiles/.
usage: ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__cluster_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg"]
def col(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name) or h.split(".", 2)[-1] == name:
            return i
    return None
for n, r in enumerate(data):
    print("--- captured launch %d ---" % n)
    for k in KEYS:
        i = col(k)
        if i is not None and i < len(r) and r[i] != "":
            print("  %-70s %s %s" % (k, r[i], units[i]))
    stalls = {}
    for i, h in enumerate(hdr):
        if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls[h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = float(r[i])
            except ValueError:
                pass
    tot = sum(stalls.values())
    if tot > 0:
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:7]
        print("  warp stall reasons (share): " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in top))
