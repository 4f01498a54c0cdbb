"""Summarise an `ncu --set full` report per launch: duration, DRAM traffic, unit utilisations, occupancy, top stall reasons.

usage: python tools/ncu_summary.py <report.ncu-rep> [title] > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return default


def to_bytes(r, name):
    v, u = f(r, name), units[col[name]] if name in col else "byte"
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def to_us(r, name):
    v, u = f(r, name), units[col[name]]
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
print("# %s\n" % title)
print("`ncu --set full --clock-control none` (one row per launch; times under the profiler are serialised and cold-cache).\n")
print("| kernel | grid x block | regs | smem KB | us | DRAM MB (r+w) | DRAM % | L2 % | L1/TEX % | SM % | tensor % | warps active % | issue active % | top stalls (warps per issue) |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("frx::", "").replace("void ", "")
    stalls = sorted(((f(r, c), c.split("stalled_")[1].split("_per_issue")[0]) for c in stall_cols), reverse=True)[:4]
    tensor = max(f(r, "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
                 f(r, "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active"),
                 f(r, "sm__pipe_tensor_subpipe_tcgen05_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_tensor_subpipe_tcgen05_cycles_active.avg.pct_of_peak_sustained_active" in col else 0.0,
                 f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else 0.0)
    print("| `%s` | %s x %s | %d | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s |" % (
        name, r[col["Grid Size"]] if "Grid Size" in col else int(f(r, "launch__grid_size")),
        r[col["Block Size"]] if "Block Size" in col else int(f(r, "launch__block_size")),
        f(r, "launch__registers_per_thread"),
        (to_bytes(r, "launch__shared_mem_per_block_dynamic") + to_bytes(r, "launch__shared_mem_per_block_static")) / 1e3,
        to_us(r, "gpu__time_duration.sum"),
        (to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")) / 1e6,
        f(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), tensor,
        f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ", ".join("%s %.2f" % (n, v) for v, n in stalls)))
