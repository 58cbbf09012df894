"""Summarise an .ncu-rep (read offline with `ncu -i`): one block per kernel launch with the metrics that matter here."""
import csv, subprocess, sys, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [
 ("time", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
 ("smem_dyn", "launch__shared_mem_per_block_dynamic"), ("waves/SM", "launch__waves_per_multiprocessor"),
 ("occ_limit_regs", "launch__occupancy_limit_registers"), ("occ_limit_smem", "launch__occupancy_limit_shared_mem"),
 ("warps_active%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("sm_thr%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
 ("issue_active%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("inst", "smsp__inst_executed.sum"),
 ("fma_pipe%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
 ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"), ("dram_thr%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
 ("l2_thr%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("l1_thr%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
 ("lts_bytes", "lts__t_bytes.sum"), ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
 ("smem_bank_conf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
 ("cycles_active", "smsp__cycles_active.avg"), ("cycles_elapsed", "sm__cycles_elapsed.max"),
 ("atom_red_inst", "smsp__inst_executed_op_global_red.sum"), ("lsu_mem_global_op_red", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum"),
]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void <unnamed>::", "")
    print("==== %s  (id %s)" % (name[:90], r[idx["ID"]]))
    for label, key in want:
        if key in idx:
            print("   %-18s %s %s" % (label, r[idx[key]], units[idx[key]]))
    st = sorted(((float(r[idx[h]] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls), reverse=True)
    print("   stalls: " + "  ".join("%s %.2f" % (n, v) for v, n in st[:7]))
