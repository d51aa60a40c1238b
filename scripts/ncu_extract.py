#!/usr/bin/env python
"""Compact extract of an `ncu --page raw --csv` export: one row per kernel with the metrics the profiles/ summaries quote.
Usage: python scripts/ncu_extract.py gpurun_out/r2a_ncu_raw_tc.csv > profiles/r2a_ncu_full_tc_kernels.csv"""
import csv, sys
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
cols = []
for w in WANT:
    for i, h in enumerate(hdr):
        if h.endswith(w):
            cols.append((w, i)); break
out = csv.writer(sys.stdout)
out.writerow(["Kernel Name"] + ["%s [%s]" % (w, units[i]) for w, i in cols])
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    out.writerow([r[ik][:100]] + [r[i] for _, i in cols])
