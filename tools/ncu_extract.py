"""Here (no GPU): turn the .ncu-rep files gpurun brought back into the committed summaries under profiles/.

    python tools/ncu_extract.py gpurun_out/r01_prof_persistent_c4.ncu-rep --pivots 8 --workload C4
    python tools/ncu_extract.py gpurun_out/r01_prof_phases_c3.ncu-rep

Writes profiles/<stem>_summary.txt (one block per profiled launch) and, with --workload, updates
profiles/r02_traffic.json (dram bytes per pivot of the persistent kernel — bench.py's roofline.traffic).
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}


def load(rep):
    if rep.endswith(".csv"):                       # raw page already exported on the GPU box (tools/profile_box.sh)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--pivots", type=int, default=0, help="pivots per profiled launch (persistent kernel)")
    ap.add_argument("--workload", default="", help="C3 | C4: update profiles/r02_traffic.json")
    ap.add_argument("--m", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    a = ap.parse_args()
    hdr, units, data = load(a.rep)
    col = {h: i for i, h in enumerate(hdr)}
    stem = os.path.splitext(os.path.basename(a.rep))[0]
    if stem.endswith("_raw"):
        stem = stem[:-4]
    lines = [f"# {stem}: ncu --set full --clock-control none, one block per profiled launch (values as ncu reports them)"]
    traffic = None
    for r in data:
        name = r[col["Kernel Name"]]
        lines.append(f"\n{name}   grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        vals = {}
        for key, label in KEYS:
            if key in col and r[col[key]] != "":
                v, u = r[col[key]], units[col[key]]
                lines.append(f"  {label:<34s} {v} {u}")
                try:
                    vals[key] = float(v.replace(",", "")) * SCALE.get(u, 1)
                except ValueError:
                    pass
        if "dram__bytes_read.sum" in vals and "dram__bytes_write.sum" in vals and "gpu__time_duration.sum" in vals:
            tot = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
            dur = vals["gpu__time_duration.sum"]
            lines.append(f"  {'dram read + write':<34s} {tot / 1e9:.4f} GB  -> {tot / dur / 1e9:.1f} GB/s over the profiled duration")
            if a.pivots and ("simplex_" in name):
                lines.append(f"  {'per pivot (' + str(a.pivots) + ' pivots in this launch)':<34s} {tot / a.pivots / 1e9:.4f} GB dram, "
                             f"{dur / a.pivots * 1e6:.1f} us (under the profiler: serialised replay passes, not a bench number)")
                traffic = {"dram_bytes_per_pivot": tot / a.pivots, "dram_read_per_pivot": vals["dram__bytes_read.sum"] / a.pivots,
                           "dram_write_per_pivot": vals["dram__bytes_write.sum"] / a.pivots, "pivots_in_capture": a.pivots,
                           "source": f"profiles/{stem}_summary.txt (ncu --set full, launch 3 of tools/prof_target.py)"}
    path = os.path.join(ROOT, "profiles", stem + "_summary.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if a.workload and traffic:
        tj = os.path.join(ROOT, "profiles", "r02_traffic.json")
        try:
            with open(tj) as f:
                allt = json.load(f)
        except Exception:
            allt = {}
        allt[a.workload] = traffic
        with open(tj, "w") as f:
            json.dump(allt, f, indent=1)
        print("updated", tj)


if __name__ == "__main__":
    main()
