"""Print the key raw metrics of every kernel in an .ncu-rep (exported with `ncu -i rep --page raw --csv`).
Usage: python tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__thread_inst_executed.sum"]
stall = [n for n in h if "issue_stalled" in n and n.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("=====", r[h.index("Kernel Name")][:90])
    for n in want:
        if n in h:
            print(f"  {n:75s} {r[h.index(n)]}")
    st = sorted(((float(r[h.index(n)] or 0), n) for n in stall), reverse=True)[:6]
    for v, n in st:
        print(f"  stall {n.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.2f}")
