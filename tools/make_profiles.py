#!/usr/bin/env python
"""Turn the raw gpurun_out/ captures of a round into the tracked summaries under profiles/."""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_atom.sum"]
MULT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def ncu_summary(rep, out_csv):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    with open(out_csv, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        w.writerow(["kernel", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "", ""])
        for k in KEEP:
            if k in hdr:
                w.writerow([k, r[hdr.index(k)], units[hdr.index(k)]])
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                w.writerow([h, r[i], units[i]])

    def val(k):
        return float(r[hdr.index(k)].replace(",", "")) * MULT[units[hdr.index(k)]]
    return val("dram__bytes_read.sum") + val("dram__bytes_write.sum")


def launch_summary(src, dst):
    shutil.copy(src, dst)
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > mv:
            d[r[kn].split("(")[0]].append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    return [(k, len(v), sum(v) / len(v) / 1e3, sum(v) / tot) for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
    traffic = {}
    traffic["heptagram"] = ncu_summary(os.path.join(G, "prof_tiles_%s.ncu-rep" % rnd), os.path.join(P, "%s_raster_tiles_heptagram_ncu.csv" % rnd))
    if os.path.exists(os.path.join(G, "prof_tiles_rgba_%s.ncu-rep" % rnd)):
        traffic["heptagram_rgba8p"] = ncu_summary(os.path.join(G, "prof_tiles_rgba_%s.ncu-rep" % rnd), os.path.join(P, "%s_raster_tiles_heptagram_rgba8p_ncu.csv" % rnd))
    if os.path.exists(os.path.join(G, "prof_tiles_b512_%s.ncu-rep" % rnd)):
        traffic["batch512"] = ncu_summary(os.path.join(G, "prof_tiles_b512_%s.ncu-rep" % rnd), os.path.join(P, "%s_raster_tiles_batch512_ncu.csv" % rnd))
    for wl in ("strokes4k", "fishy256"):
        rep = os.path.join(G, "prof_tiles_%s_%s.ncu-rep" % (wl, rnd))
        if os.path.exists(rep):
            traffic[wl + ("_rgba8p" if wl == "strokes4k" else "")] = ncu_summary(rep, os.path.join(P, "%s_raster_tiles_%s_ncu.csv" % (rnd, wl)))
    traffic["source"] = "profiles/%s_raster_tiles_*_ncu.csv: dram__bytes_read.sum + dram__bytes_write.sum of one raster_tiles launch (ncu --set full)" % rnd
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    for name in ("launches_%s.csv" % rnd, "launches_b512_%s.csv" % rnd):
        if os.path.exists(os.path.join(G, name)):
            tag = "batch512" if "b512" in name else "heptagram"
            for k, n, avg, share in launch_summary(os.path.join(G, name), os.path.join(P, "%s_launches_%s.csv" % (rnd, tag))):
                print("%-10s %-42s n=%3d avg=%9.1f us share=%.4f" % (tag, k[:42], n, avg, share))
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
