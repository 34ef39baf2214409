"""Summarise an .ncu-rep (one kernel) into profiles/<name>.md and <name>.json.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_lcs_tile [kernel-index]"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
def to_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    return float(v.replace(",", "")) * f
sel = {k: {"value": m[k][0], "unit": m[k][1]} for k in keys if k in m}
dram = None
if "dram__bytes_read.sum" in m:
    dram = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
json.dump({"source": rep, "dram_bytes_per_launch": dram, "metrics": sel}, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as fh:
    fh.write(f"# ncu --set full summary: {m.get('Kernel Name', ('?',))[0][:80]}\n\nsource: `{rep}` (kernel index {idx})\n\n| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in m:
            fh.write(f"| {k} | {m[k][0][:90]} | {m[k][1]} |\n")
    fh.write(f"\nDRAM traffic per launch (read+write): {dram:.0f} bytes\n" if dram is not None else "")
print(open(out + ".md").read())
