"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
args = ["ncu", "-i", rep, "--page", "raw", "--csv"]
if len(sys.argv) > 2:
    args += ["--kernel-name", "regex:" + sys.argv[2]]
txt = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
    "sm__inst_executed_pipe_uniform.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("=" * 30, r[ix["Kernel Name"]][:90])
    for w in WANT:
        if w in ix:
            print(f"{w:75s} {r[ix[w]]:>20s} {units[ix[w]]}")
    st = sorted(((float(r[ix[h]] or 0), h) for h in STALL), reverse=True)[:7]
    for v, h in st:
        print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:8.3f}")
