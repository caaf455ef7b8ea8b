#!/usr/bin/env python3
"""`ncu -i X.ncu-rep --page raw --csv | python tools/ncu_to_json.py "<how it was captured>" > profiles/r2_ncu_summary.json`
Per kernel (the LONGEST launch of each name): duration, executed warp instructions, FP64 / XU pipe and issue utilisation,
warps active, DRAM bytes, launch shape.  bench.py quotes these as the kernels' executed work."""
import csv, json, re, sys
rows = list(csv.reader(sys.stdin))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = {"duration_us": ("gpu__time_duration.sum", 1e-3), "inst_executed": ("smsp__inst_executed.sum", 1),
        "fp64_pipe_pct": ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1),
        "xu_pipe_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
        "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        "dram_read_bytes": ("dram__bytes_read.sum", 1), "dram_write_bytes": ("dram__bytes_write.sum", 1),
        "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        "registers": ("launch__registers_per_thread", 1), "grid": ("launch__grid_size", 1), "block": ("launch__block_size", 1),
        "threads_per_inst": ("smsp__thread_inst_executed_per_inst_executed.ratio", 1)}
units = rows[1]
out = {"_how": sys.argv[1] if len(sys.argv) > 1 else "ncu --set full --clock-control none"}
for r in data:
    name = re.sub(r"^void ", "", r[col["Kernel Name"]])
    name = re.sub(r"\(.*$", "", name).replace("gort::", "")
    d = {}
    for k, (m, f) in want.items():
        if m in col and r[col[m]] not in ("", "n/a"):
            v = float(r[col[m]].replace(",", ""))
            u = units[col[m]]
            if k == "duration_us":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            elif k.startswith("dram_") and k.endswith("bytes"):
                v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            else:
                v = v * f
            d[k] = round(v, 3)
    if name.startswith("at::"):
        continue                                   # torch's own fill / copy kernels are not ours
    base = re.sub(r"<.*$", "", name)
    for key in {name, base}:
        if key not in out or d.get("duration_us", 0) > out[key].get("duration_us", 0):
            out[key] = d
print(json.dumps(out, indent=1))
