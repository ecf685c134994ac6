"""Key metrics of an `ncu --set full` report (read on the CPU box): python tools/ncu_metrics.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w} [{units[i]}]: {[r[i] for r in data]}")
print("-- warp stall reasons (cycles per issued instruction) --")
st = []
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False and h.endswith(".ratio"):
        name = h.split("issue_stalled_")[1].replace(".ratio", "")
        try:
            st.append((float(data[0][i]), name, [r[i] for r in data]))
        except ValueError:
            pass
for v, name, vals in sorted(st, reverse=True)[:12]:
    print(f"  {name}: {vals}")
