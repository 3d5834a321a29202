"""Key counters of every launch in an .ncu-rep: python tools/ncu_keys.py report.ncu-rep [kernel-regex]"""
import csv, subprocess, sys
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
if len(sys.argv) > 2:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
KEYS = [("gpu__time_duration.sum", "dur_us"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_inst%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefronts%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ_regs"),
        ("launch__occupancy_limit_shared_mem", "occ_smem"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1_ld_sectors"),
        ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("sm__sass_inst_executed_op_local_ld.sum", "local_ld"), ("smsp__inst_executed_op_local_st.sum", "local_st")]
stall = [k for k in h if k.startswith("smsp__average_warp") and k.endswith("_per_issue_active.ratio") or
         (k.startswith("smsp__average_warps_issue_stalled") and k.endswith(".ratio"))]
for r in rows[2:]:
    d = dict(zip(h, r))
    print("==", d.get("Kernel Name", "")[:110])
    for k, nm in KEYS:
        if k in d:
            print(f"  {nm:18s} {d[k]}")
    st = sorted(((float(d[k]), k) for k in stall if d.get(k) not in (None, "", "n/a")), reverse=True)[:7]
    for v, k in st:
        print(f"  stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.2f}")
