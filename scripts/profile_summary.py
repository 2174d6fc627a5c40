"""Turns the ncu captures gpurun brought back (gpurun_out/<tag>_full.ncu-rep, <tag>_launches.csv) into the tracked
summaries under profiles/: <tag>_ncu_summary.txt, <tag>_launches.csv and profiles/traffic.json (DRAM bytes per launch of
each kernel, which bench.py reports as roofline.traffic).   usage: python scripts/profile_summary.py <tag> "<one-line note>"
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rep = os.path.join(ROOT, "gpurun_out", tag + "_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
lines = ["%s: ncu --set full --clock-control none, python bench.py --steps 1 --warmup 0 (n=203 aniso, fp fast). %s" % (tag, note)]
traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    vals = []
    dram = 0.0
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals.append("%s=%s %s" % (k, r[i], units[i]))
            if k.startswith("dram__bytes"):
                dram += float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    stalls = []
    for i, h in enumerate(hdr):
        if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    lines.append("%s\n    %s\n    stalls per issue: %s" % (short, ", ".join(vals), ", ".join("%s %.2f" % (h, v) for v, h in stalls[:6])))
    def pct(key):
        try:
            return float(r[hdr.index(key)].replace(",", ""))
        except (ValueError, IndexError):
            return None
    # the units that bound these kernels besides HBM: reported by bench.py next to the HBM fraction
    traffic[short] = {"dram_bytes_per_launch": dram, "source": "profiles/%s_ncu_summary.txt" % tag,
                      "fp64_pipe_pct": pct("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                      "l1_data_pipe_pct": pct("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                      "dram_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                      "warps_active_pct": pct("sm__warps_active.avg.pct_of_peak_sustained_active")}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
launch_src = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
if os.path.exists(launch_src):
    shutil.copy(launch_src, os.path.join(ROOT, "profiles", tag + "_launches.csv"))
    lr = [x for x in csv.reader(open(launch_src)) if len(x) > 10]
    h2 = lr[0]
    agg = collections.OrderedDict()
    for x in lr[1:]:
        agg.setdefault(x[h2.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), []).append(
            float(x[h2.index("Metric Value")].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines.append("launch list (%s_launches.csv, gpu__time_duration.sum, serialised cold-cache launches): share of all kernel time" % tag)
    for k, v in agg.items():
        lines.append("    %-28s launches=%d mean=%.1f us share=%.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
print("\n".join(lines))
