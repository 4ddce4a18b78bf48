"""Dev tool: turn an .ncu-rep of `tools/ncu_target.py` into the committed summaries: profiles/<name>_raw.txt (key raw metrics and stall
reasons per kernel) and, with --traffic <json>, the DRAM-bytes record `bench.py` reads for `roofline.traffic` (keyed by the sha256 of
the kernel's source file).  usage: ncu_summary.py <rep> <out_raw.txt> "<header line>" [--traffic <json> <source.cu> "<workload>"]"""
import csv, hashlib, json, subprocess, sys
rep, out_txt, header = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
lines = [header, ""]
first = None
for vals in rows[2:]:
    if len(vals) != len(hdr): continue
    get = lambda k: vals[hdr.index(k)] if k in hdr else ""
    if first is None: first = vals
    lines.append(f"== {get('Kernel Name')}  (grid {get('launch__grid_size')} x {get('launch__block_size')})")
    for k in want:
        if k in hdr: lines.append(f"{k:92s} {vals[hdr.index(k)]} {units[hdr.index(k)]}")
    for i, k in enumerate(hdr):
        if "issue_stalled" in k and "per_issue_active" in k:
            try:
                if float(vals[i] or 0) > 0.05: lines.append(f"{k:92s} {vals[i]}")
            except ValueError: pass
    lines.append("")
open(out_txt, "w").write("\n".join(lines))
print("\n".join(lines[:12]))
if "--traffic" in sys.argv:
    i = sys.argv.index("--traffic")
    path, src, workload = sys.argv[i + 1:i + 4]
    g = lambda k: first[hdr.index(k)]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    b = lambda k: int(round(float(g(k)) * scale[units[hdr.index(k)]]))
    tu = units[hdr.index("gpu__time_duration.sum")]
    t = float(g("gpu__time_duration.sum")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1e-6)
    rec = {"kernel": g("Kernel Name"), "workload": workload, "dram_bytes_read": b("dram__bytes_read.sum"),
           "dram_bytes_write": b("dram__bytes_write.sum"), "gpu_time_ms": t,
           "source_sha16": hashlib.sha256(open(src, "rb").read()).hexdigest()[:16],
           "capture": f"{rep} (ncu --set full --clock-control none)"}
    json.dump(rec, open(path, "w"), indent=1)
    print(rec)
