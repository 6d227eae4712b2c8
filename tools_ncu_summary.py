#!/usr/bin/env python3
"""tools_ncu_summary.py REPORT.ncu-rep OUT.txt [TRAFFIC.json PARTICLES] — condenses an `ncu --set full` report into the text summary kept under
profiles/: per launch the duration, DRAM bytes, L2/L1/SM throughput, shared-memory wavefronts and bank conflicts,
occupancy, registers, and the warp-stall sample breakdown of each distinct kernel."""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(h)}
COLS = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"), ("smsp__inst_executed.sum", "warp_inst")]


def val(row, name):
    i = ix.get(name)
    if i is None:
        return None
    try:
        v = float(row[i].replace(",", ""))
    except ValueError:
        return None
    u = units[i]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    return v


with open(out, "w") as f:
    f.write("# condensed from %s (ncu --set full --clock-control none); one line per profiled launch\n" % rep.split("/")[-1])
    f.write("%-44s " % "kernel" + " ".join("%14s" % c[1] for c in COLS) + "\n")
    seen = {}
    for r in data:
        name = r[ix["Kernel Name"]]
        short = name.replace("vfd::", "").split("(")[0].replace("void ", "")
        f.write("%-44s " % short[:44] + " ".join("%14s" % ("-" if val(r, c[0]) is None else ("%.4g" % val(r, c[0]))) for c in COLS) + "\n")
        seen.setdefault(short, r)
    f.write("\n# warp-stall sampling (pc samples per reason) of the first launch of each kernel\n")
    for short, r in seen.items():
        st = []
        for n in h:
            if n.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in n:
                try:
                    v = float(r[ix[n]])
                except ValueError:
                    continue
                if v > 0:
                    st.append((v, n.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        tot = sum(v for v, _ in st) or 1.0
        f.write("%-44s " % short[:44] + "  ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in sorted(st, reverse=True)[:7]) + "\n")
print(open(out).read())

# optional: DRAM bytes per launch per bench.py kernel class (mean over the captured launches) -> profiles/ncu_traffic.json
if len(sys.argv) > 4:
    import json
    CLASS = [("k_visc_matvec_pipe<0>", "visc_matvec"), ("k_visc_matvec_pipe<1>", "visc_matvec0"), ("k_visc_update", "visc_update"),
             ("k_visc_direction", "visc_direction"), ("k_visc_setup", "visc_setup"), ("k_density_factor", "density_factor"),
             ("k_solve_iteration<1>", "div_solve"), ("k_solve_iteration<0>", "press_solve"),
             ("k_source<1>", "div_source"), ("k_source<0>", "press_source"),
             ("k_pressure_accel<0>", "div_accel"), ("k_pressure_accel<1>", "div_finish"),
             ("k_pressure_accel<2>", "press_accel"), ("k_pressure_accel<3>", "press_finish"),
             ("k_st_classify", "st_classify"), ("k_st_smooth", "st_smooth"), ("k_build_list", "search_build_list")]
    acc = {}
    for r in data:
        name = r[ix["Kernel Name"]].replace("(bool)", "").replace("(int)", "")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        if rd is None or wr is None:
            continue
        for pat, cls in CLASS:
            if pat in name:
                acc.setdefault(cls, []).append(rd + wr)
                break
    with open(sys.argv[3], "w") as f:
        json.dump({"source": rep.split("/")[-1] + " (ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches)",
                   "particles": int(sys.argv[4]), "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in acc.items()}}, f, indent=1)
