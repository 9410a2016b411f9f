"""Summarise an ncu --set full report (read here, no GPU needed): one row per profiled launch with the metrics the
roofline discussion uses.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.md "title" """
import csv, io, json, subprocess, sys, re

rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
     ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex %"),
     ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu wavefronts %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
     ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
     ("l1tex__t_sector_hit_rate.pct", "l1 hit %"), ("lts__t_sector_hit_rate.pct", "l2 hit %"),
     ("launch__registers_per_thread", "regs")]
idx = [(hdr.index(m), lab) for m, lab in M if m in hdr]
ik = hdr.index("Kernel Name")
traffic = {}
l1busy = {}
with open(out, "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` (1x B200, under gpurun); read with `ncu -i ... --page raw --csv`.\n\n")
    f.write("| kernel | " + " | ".join(l for _, l in idx) + " |\n|---|" + "---:|" * len(idx) + "\n")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("dfr::", "")
        vals = []
        for i, lab in idx:
            v = r[i]
            try:
                x = float(v.replace(",", ""))
                v = f"{x:.1f}" if abs(x) < 1e6 else f"{x:.3g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[i]}".strip() if lab in ("time", "dram rd", "dram wr") else v)
        f.write(f"| {name} | " + " | ".join(vals) + " |\n")
        try:
            def val(m):
                i = hdr.index(m)
                x = float(r[i].replace(",", ""))
                u = units[i].lower()
                return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            base = name.split("<")[0]
            traffic.setdefault(base, []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
            l1busy.setdefault(base, []).append(float(r[hdr.index("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")].replace(",", "")) / 100.0)
        except Exception:
            pass
print(open(out).read())
if len(sys.argv) > 4:
    d = {k: sum(v) / len(v) for k, v in traffic.items()}
    # busiest unit of the list kernels: the L1 data stage (share of its peak wavefront rate, mean over the profiled launches)
    d["l1_data_stage_busy"] = {k: sum(v) / len(v) for k, v in l1busy.items()}
    json.dump(d, open(sys.argv[4], "w"), indent=1)
