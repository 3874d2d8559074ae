"""Summarise an .ncu-rep: headline metrics + stall hot spots per SASS line (reads `ncu -i` CSV pages)."""
import csv, subprocess, sys, io

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_xu.sum",
        "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_uniform.sum", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h} = {vals[i]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print("samples", tot, "inst", sum(int(r[ix["Instructions Executed"]]) for r in data))
print(sorted(agg.items(), key=lambda kv: -kv[1])[:9])
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:topn]
for i in sorted(top):
    r = data[i]
    st = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0}
    print(i, r[ix["Source"]].strip()[:58].ljust(58), r[ix["# Samples"]], r[ix["Instructions Executed"]], st)
