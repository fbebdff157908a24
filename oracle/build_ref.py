"""Builds the reference's OWN CUDA extension (taichi_splatting/cuda_lib: full_cumsum, radix_sort_pairs,
segmented_sort_pairs) for sm_100a into oracle/_ref/, from the sources where they lie under /root/reference.

TEST INFRASTRUCTURE ONLY: the built module is the real reference for rows R4 (scan + total) and R6 (radix sort of
(key, value) pairs); the GPU tests compare gs_tile_scan / gs_sort_pairs with it and profiles/ref_cuda_lib_bar.py
times both.  Everything else on the render path is Taichi and cannot be built here.  oracle/_ref/ is git-ignored
(built artefact) but travels to the GPU box with the snapshot.  Nothing is copied from the reference tree.

  python oracle/build_ref.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("GS_REFERENCE_ROOT", "/root/reference") + "/taichi_splatting/cuda_lib"
NAME = "ref_cuda_lib"
SOURCES = ["full_cumsum.cu", "radix_sort_pairs.cu", "segmented_sort_pairs.cu", "module.cpp"]


def available() -> bool:
  return all(os.path.isfile(os.path.join(REF, s)) for s in SOURCES)


def built_path():
  p = os.path.join(OUT, NAME + ".so")
  return p if os.path.isfile(p) else None


def build(verbose: bool = False):
  """Compiles (ninja, a few minutes the first time) and returns the path of the module, or None without the tree."""
  if built_path() is not None:
    return built_path()
  if not available():
    return None
  os.makedirs(OUT, exist_ok=True)
  os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
  os.environ.setdefault("MAX_JOBS", "4")
  from torch.utils.cpp_extension import load
  load(NAME, sources=[os.path.join(REF, s) for s in SOURCES], build_directory=OUT, verbose=verbose,
       extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], is_python_module=False)
  for f in os.listdir(OUT):   # keep the module only: the snapshot that travels to the GPU box stays small
    if not f.endswith(".so"):
      os.remove(os.path.join(OUT, f))
  return built_path()


def load_module():
  """Imports the built module (GPU box: the prebuilt oracle/_ref/ref_cuda_lib.so) or returns None."""
  path = built_path()
  if path is None:
    return None
  import importlib.util
  import torch  # noqa: F401  (the extension links against libtorch)
  spec = importlib.util.spec_from_file_location(NAME, path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


if __name__ == "__main__":
  print(build(verbose="-v" in sys.argv))
