"""Condense an ncu report into the JSON / text summaries kept under profiles/.
usage: ncu_summary.py <report.ncu-rep> <out_prefix>   (writes <out_prefix>_ncu_full.json, _sass_hist.txt, _stalls.txt)"""
import csv, io, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "sm__cycles_elapsed.max",
        "gpc__cycles_elapsed.avg.per_second", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "smsp__sass_branch_targets_threads_divergent.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"]
d = {}
for k in keep:
    if k in hdr:
        i = hdr.index(k)
        d[k] = {"value": vals[i], "unit": units[i]}
for h_ in hdr:
    if "warp_issue_stalled" in h_ and h_.endswith("_per_warp_active.pct"):
        d[h_] = {"value": vals[hdr.index(h_)], "unit": "%"}
json.dump(d, open(out + "_ncu_full.json", "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
tmp = out + "_src_sass.tmp.csv"
open(tmp, "w").write(src)
open(out + "_sass_hist.txt", "w").write(subprocess.run([sys.executable, os.path.join(here, "sass_hist.py"), tmp, "40"], capture_output=True, text=True).stdout)
open(out + "_stalls.txt", "w").write(subprocess.run([sys.executable, os.path.join(here, "sass_stalls.py"), tmp], capture_output=True, text=True).stdout)
os.remove(tmp)
rd = float(d["dram__bytes_read.sum"]["value"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"]["unit"]]
wr = float(d["dram__bytes_write.sum"]["value"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"]["unit"]]
print("traffic bytes per launch:", rd + wr)
for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"):
    print(k, d[k]["value"])
