"""Builds libgsplat_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

  python -m taichi_splatting_b200.build [--force] [--verbose]

One nvcc invocation per translation unit (parallel), then one link.  No torch headers are involved:
the library is plain CUDA behind `extern "C"` (include/gsplat_b200.h).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
VARIANT = os.environ.get("GS_BUILD_VARIANT", "")          # experiments: alternate .so + extra -D flags
OBJ = PKG / "csrc" / ("build" + VARIANT)
LIB = PKG / f"libgsplat_b200{VARIANT}.so"
SOURCES = ["api.cu", "projection.cu", "sh.cu", "mapper.cu", "raster_generic.cu", "raster_digest.cu", "raster_pack.cu", "raster_fwd.cu", "raster_fwd_bulk.cu", "raster_bwd.cu", "raster_bwd_t.cu", "render.cu", "optim.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr",
] + os.environ.get("GS_BUILD_FLAGS", "").split()


def _nvcc() -> str:
  for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return "nvcc"


def _deps_mtime() -> float:
  files = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "gsplat_b200.h"]
  return max(f.stat().st_mtime for f in files)


def build(force: bool = False, verbose: bool = False) -> Path:
  OBJ.mkdir(parents=True, exist_ok=True)
  hdr_mtime = _deps_mtime()
  jobs = []
  for src in SOURCES:
    s, o = CSRC / src, OBJ / (src + ".o")
    if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, hdr_mtime):
      jobs.append((s, o))

  def compile_one(job):
    s, o = job
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
    if verbose:
      cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    return s.name, r

  if jobs:
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
      for name, r in ex.map(compile_one, jobs):
        if verbose or r.returncode != 0:
          sys.stderr.write(f"--- {name}\n{r.stdout}{r.stderr}\n")
        if r.returncode != 0:
          raise RuntimeError(f"nvcc failed for {name}")
  objs = [str(OBJ / (s + ".o")) for s in SOURCES]
  if jobs or force or not LIB.exists():
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-ccbin", "/usr/bin/g++", "-cudart", "shared"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      sys.stderr.write(r.stdout + r.stderr)
      raise RuntimeError("link failed")
  return LIB


if __name__ == "__main__":
  ap = argparse.ArgumentParser()
  ap.add_argument("--force", action="store_true")
  ap.add_argument("--verbose", action="store_true")
  a = ap.parse_args()
  print(build(a.force, a.verbose))
