"""Per-kernel DRAM traffic, duration and pipe utilisation from an `ncu -i <rep> --page raw --csv` dump.
Usage: python tools/ncu_traffic.py gpurun_out/<tag>_accumulate_ncu_raw.csv <tag> > profiles/accumulate_traffic.json
The capture holds the kernels of ONE bucket-accumulate pass (every affine pair-tree round + the XYZZ slices); the JSON
sums them (`dram_bytes_per_launch` = per pass = per MSM) and keeps the per-kernel rows."""
import csv, json, re, sys

path, tag = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path, newline="") if l.startswith('"')]
rd = list(csv.reader(lines))
hdr, units, body = rd[0], rd[1], rd[2:]
col = {n: i for i, n in enumerate(hdr)}


def num(row, name, default=None):
    i = col.get(name)
    if i is None or row[i] in ("", "n/a"):
        return default
    v = float(row[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1,
             "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1}.get(u, 1)
    return v * scale


kern = []
for r in body:
    name = re.sub(r"\(.*$", "", r[col["Kernel Name"]])
    kern.append({
        "kernel": name,
        "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
        "duration_ms": (num(r, "gpu__time_duration.sum") or 0) * 1e3,
        "dram_read": num(r, "dram__bytes_read.sum"), "dram_write": num(r, "dram__bytes_write.sum"),
        "fmaheavy_cycles_pct": num(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "fma_inst_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "alu_inst_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "ipc": num(r, "sm__inst_executed.avg.per_cycle_active"),
        "registers": num(r, "launch__registers_per_thread"),
        "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    })
tot_r = sum(k["dram_read"] or 0 for k in kern)
tot_w = sum(k["dram_write"] or 0 for k in kern)
print(json.dumps({
    "round": tag, "pass": "bucket accumulate = affine pair-tree rounds + XYZZ slices (one MSM, 2^20 pairs -> 2^21 after the split, c = 16, W = 8)",
    "dram_bytes_per_launch": tot_r + tot_w, "dram_read": tot_r, "dram_write": tot_w,
    "duration_ms_under_ncu": sum(k["duration_ms"] for k in kern),
    "source": "ncu --set full --clock-control none, raw page (dram__bytes_read.sum + dram__bytes_write.sum per kernel, summed over the pass)",
    "kernels": kern}, indent=1))
