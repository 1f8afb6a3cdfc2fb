#!/usr/bin/env python
"""Turn the raw gpurun_out/ captures of a round (tools/gpu_profiles_r2.sh writes them as p2_*) into the tracked
summaries under profiles/: bench lines copied, launch lists summarised, one metric table per full ncu capture,
and profiles/traffic.json (DRAM bytes per launch of the dominant kernel, read by bench.py)."""
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
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_atom.sum"]
MULT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def ncu_summary(raw_csv, out_csv):
    """raw_csv: `ncu -i report --page raw --csv` of one launch (exported on the GPU box: the reports are too big to bring back)."""
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, r = rows[0], rows[1], rows[2]
    with open(out_csv, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        w.writerow(["kernel", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "", ""])
        for k in KEEP:
            if k in hdr:
                w.writerow([k, r[hdr.index(k)], units[hdr.index(k)]])
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                w.writerow([h, r[i], units[i]])

    def val(k):
        return float(r[hdr.index(k)].replace(",", "")) * MULT[units[hdr.index(k)]]
    return val("dram__bytes_read.sum") + val("dram__bytes_write.sum")


def launch_summary(src, dst):
    shutil.copy(src, dst)
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > mv:
            v = float(r[mv].replace(",", ""))
            v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
            d[r[kn].split("(")[0]].append(v)
    tot = sum(sum(v) for v in d.values())
    return [(k, len(v), sum(v) / len(v), sum(v) / tot) for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
    pre = "p2_"
    traffic = {}
    for wl, key in (("heptagram", "heptagram"), ("heptagram_rgba8p", "heptagram_rgba8p"), ("batch512", "batch512"), ("bigraster", "bigraster"),
                    ("strokes4k", "strokes4k_rgba8p"), ("fishy256", "fishy256"), ("small", None), ("stroke_segments", None)):
        rep = os.path.join(G, "%sraw_%s.csv" % (pre, wl))
        if os.path.exists(rep) and os.path.getsize(rep) > 0:
            t = ncu_summary(rep, os.path.join(P, "%s_%s_ncu.csv" % (rnd, wl)))
            if key:
                traffic[key] = t
    old = {}
    try:
        old = json.load(open(os.path.join(P, "traffic.json")))
    except Exception:
        pass
    for k, v in old.items():
        if k not in traffic and k != "source":
            traffic[k] = v
    traffic["source"] = "profiles/%s_*_ncu.csv: dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant tile kernel (ncu --set full)" % rnd
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    with open(os.path.join(P, "%s_launch_shares.txt" % rnd), "w") as f:
        for wl in ("heptagram", "batch512", "bigraster", "strokes4k", "fishy256", "stroke"):
            src = os.path.join(G, "%slaunches_%s.csv" % (pre, wl))
            if os.path.exists(src):
                for k, n, avg, share in launch_summary(src, os.path.join(P, "%s_launches_%s.csv" % (rnd, wl))):
                    line = "%-10s %-44s n=%3d avg=%10.1f us share=%.4f" % (wl, k[:44], n, avg, share)
                    print(line)
                    f.write(line + "\n")
    for name in os.listdir(G):
        if name.startswith(pre + "bench_") and name.endswith(".json") and os.path.getsize(os.path.join(G, name)) > 0:
            shutil.copy(os.path.join(G, name), os.path.join(P, rnd + "_" + name[len(pre):]))
    if os.path.exists(os.path.join(G, pre + "stroke_probe.txt")):
        shutil.copy(os.path.join(G, pre + "stroke_probe.txt"), os.path.join(P, rnd + "_stroke_probe.txt"))
    if os.path.exists(os.path.join(G, pre + "smi.csv")):
        shutil.copy(os.path.join(G, pre + "smi.csv"), os.path.join(P, rnd + "_nvidia_smi_after.csv"))
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
