"""Turns gpurun_out ncu artefacts into the committed summaries under profiles/.

  python profiles/summarize.py <tag> [steps]   (expects gpurun_out/launches_<tag>.csv and gpurun_out/raster_<tag>.ncu-rep;
                                                `steps`: the launch list covers that many steps, keep the last one)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_md = os.path.join(ROOT, "profiles", f"{tag}_summary.md")
lines = [f"# ncu summary `{tag}` (cfg3: 1 M Gaussians, 2048x2048, SH deg 3, vis + heuristics + median depth; one fwd+bwd step)", ""]

# ---- launch list: per-kernel device time (cold cache, serialised: compare shares) ----
path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if steps > 1:   # warm step only: the first steps pay module loading and cold caches
  # first kernel of a step: the compaction init of the single-pass projection (round 2), else the round-1 cull kernel
  starts = [i for i, r in enumerate(rows) if "DeviceCompactInitKernel" in r["Kernel Name"]] or \
           [i for i, r in enumerate(rows) if "project_cull" in r["Kernel Name"]]
  rows = rows[starts[-1]:]
agg, tot = {}, 0.0
for r in rows:
  name = r["Kernel Name"]
  v = float(r["Metric Value"].replace(",", ""))
  a = agg.setdefault(name, [0.0, 0])
  a[0] += v
  a[1] += 1
  tot += v
lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)", "",
          "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
  if v / tot > 0.001:
    lines.append(f"| `{k[:110]}` | {c} | {v / 1e3:.1f} | {100 * v / tot:.1f} % |")
lines += [f"| **total** | {len(rows)} | {tot / 1e3:.1f} | 100 % |", ""]

# ---- full captures of the raster kernels ----
rep = os.path.join(ROOT, "gpurun_out", f"raster_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
traffic = {}
lines += ["## `ncu --set full` captures (every hand-written kernel of the step)", ""]
for r in rr[2:]:
  if len(r) != len(hdr):
    continue
  name = r[hdr.index("Kernel Name")]
  lines += [f"### `{name[:120]}`", "", "| metric | value | unit |", "|---|---:|---|"]
  vals = {}
  for w in want:
    if w in hdr:
      vals[w] = r[hdr.index(w)]
      lines.append(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")
  lines.append("")

  def to_bytes(key):
    v, u = float(vals[key].replace(",", "")), units[hdr.index(key)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
  key = "raster_bwd_kernel" if "raster_bwd" in name else name.split("(")[0]
  traffic[key] = {"dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                  "capture": f"raster_{tag}.ncu-rep", "kernel": name[:100]}
  # what actually bounds the kernel: utilisation of the L1/shared-memory data pipe and of the issue slots
  for metric, short in (("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_data_pipe_pct"),
                        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slots_pct"),
                        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct")):
    if metric in vals:
      traffic[key][short] = float(vals[metric].replace(",", ""))
open(out_md, "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(out_md)
