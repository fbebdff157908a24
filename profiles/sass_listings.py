"""Writes the `cuobjdump -sass` listings of the bench instantiations of the raster kernels (F = 3, visibility /
heuristics / median) from the built library into profiles/<round>/sass_*.txt, each with a static opcode histogram.
Usage: python profiles/sass_listings.py [out_dir]      (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02")
lib = os.path.join(ROOT, "taichi_splatting_b200", "libgsplat_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
KERNELS = {"sass_raster_fwd_bulk.txt": "raster_fwd_bulk_kernelILi3ELb1ELb1ELi3",
           "sass_raster_bwd.txt": "raster_bwd_t_kernelILi3ELb1ELb1ELb1ELi3",
           "sass_raster_pack.txt": "raster_pack_flat_kernelILi3ELb1"}
for fname, key in KERNELS.items():
  body = next(f for f in funcs if key in f.split("\n", 1)[0])
  lines = []
  ops = collections.Counter()
  for l in body.split("\n"):
    m = re.match(r"\s+(/\*[0-9a-f]{4}\*/)\s+(.*?);", l)
    if m:
      lines.append(f"{m.group(1)} {m.group(2).strip()} ;")
      op = re.sub(r"^@!?U?P\d\s+", "", m.group(2).strip()).split()[0].split(".")[0]
      ops[op] += 1
  hist = ", ".join(f"{k} {v}" for k, v in ops.most_common(28))
  evid = ", ".join(f"{k} {ops[k]}" for k in ("SYNCS", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "SHFL", "ATOMS") if ops[k])
  with open(os.path.join(out_dir, fname), "w") as f:
    f.write(f"# cuobjdump -sass of libgsplat_b200.so (sm_100a), kernel {key} -- bench instantiation (F = 3, visibility / heuristics / median).\n")
    f.write(f"# {sum(ops.values())} instructions.  Static opcode histogram: {hist}\n")
    f.write(f"# Bulk-copy / mbarrier / packed-f32 evidence: {evid}\n")
    f.write("Function : " + body.split("\n", 1)[0].strip() + "\n")
    f.write("\n".join(lines) + "\n")
  print(fname, sum(ops.values()), evid)
