#!/usr/bin/env python
"""Reads the CSV pages scripts/ncu_tc.sh writes: key metrics per kernel (raw page) and the instructions that hold the
most stall samples (source page).  Usage: python scripts/ncu_read.py gpurun_out/r2d_p0 [--top 40] [--kernel wgrad]"""
import csv, gzip, sys, argparse
ap = argparse.ArgumentParser(); ap.add_argument("base"); ap.add_argument("--top", type=int, default=30)
ap.add_argument("--kernel", default=""); ap.add_argument("--nosrc", action="store_true")
a = ap.parse_args()
KEYS = ["gpu__time_duration.sum", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_subpipe", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__warps_active",
        "smsp__average_warp", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_op", "tensor"]
rows = list(csv.reader(open(a.base + "_raw.csv")))
hdr = rows[0]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if a.kernel and a.kernel not in name: continue
    print("==", name[:90])
    for i, h in enumerate(hdr):
        if any(k in h for k in KEYS):
            print("   %-90s %s %s" % (h, r[i], rows[1][i]))
if a.nosrc: sys.exit()
rows = list(csv.reader(gzip.open(a.base + "_source.csv.gz", "rt")))
ks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; ks.append(cur); continue
    if cur is None or not r: continue
    if r[0] == "Address": cur["hdr"] = r; continue
    cur["rows"].append(r)
for k in ks:
    if a.kernel and a.kernel not in k["name"]: continue
    h = k["hdr"]; ia = h.index("Warp Stall Sampling (All Samples)"); isrc = h.index("Source"); iex = h.index("Instructions Executed")
    tot = sum(int(r[ia]) for r in k["rows"])
    print("==", k["name"][:90], "samples", tot)
    top = sorted(enumerate(k["rows"]), key=lambda t: -int(t[1][ia]))[:a.top]
    for i, r in sorted(top):
        st = {h[j]: int(r[j]) for j in range(len(h)) if h[j].startswith("stall_") and "Not" not in h[j] and r[j] not in ("", "0")}
        s = sorted(st.items(), key=lambda t: -t[1])[:2]
        print("%5d %-72s %6s %10s %s" % (i, r[isrc].strip()[:72], r[ia], r[iex], s))
