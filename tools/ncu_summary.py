"""Summarise an .ncu-rep (run where ncu is installed): per kernel duration, pipe utilisation, stalls, DRAM bytes."""
import csv, subprocess, sys, json
rep = sys.argv[1]
if rep.endswith(".csv"):      # `ncu -i X.ncu-rep --page raw --csv` exported on the GPU box (reports can exceed the 64 MiB pull)
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "time",
    "sm__cycles_elapsed.avg": "cycles",
    "sm__inst_executed.sum.pct_of_peak_sustained_elapsed": "issue_pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active": "uniform_pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_tc_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_lsu_pct",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "smsp__inst_executed.sum": "warp_insts",
    "launch__registers_per_thread": "regs",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefront_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
}
res = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    k = {"kernel": d["Kernel Name"][:60]}
    for name, short in want.items():
        for h in hdr:
            if h.endswith(name):
                k[short] = d[h] + (" " + u[h] if u[h] else "")
                break
    st = {}
    for h in hdr:
        if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio"):
            st[h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = float(d[h] or 0)
    k["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:6])
    res.append(k)
print(json.dumps(res, indent=1))
