"""Load the reference's own pure-torch `torch_lib` as an oracle.

Works only where /root/reference exists (the build container).  taichi / tensordict / roma are
absent, so the four modules torch_lib imports from are stubbed.  Used by
tests/golden/make_golden.py (fixture generation) and by the CPU tests that cross-check
oracle/torch_ops.py when the reference tree is present.  Never used on the GPU box.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("GS_REFERENCE_ROOT", "/root/reference") + "/taichi_splatting"


def available() -> bool:
  return os.path.isfile(REF + "/torch_lib/projection.py")


def load():
  """-> (projection_module, spherical_harmonics_module) of the reference torch_lib."""
  def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m

  def _ours(k):
    return k == "taichi_splatting" or k.startswith("taichi_splatting.")

  saved = {k: v for k, v in sys.modules.items() if _ours(k)}
  try:
    _mod("taichi_splatting").__path__ = []
    _mod("taichi_splatting.perspective", CameraParams=object)
    _mod("taichi_splatting.data_types", Gaussians3D=object, RasterConfig=object)
    _mod("taichi_splatting.taichi_queue", queued=lambda f: f)
    _mod("taichi_splatting.torch_lib").__path__ = [REF + "/torch_lib"]

    def _load(name, file):
      spec = importlib.util.spec_from_file_location(name, file)
      m = importlib.util.module_from_spec(spec)
      sys.modules[name] = m
      spec.loader.exec_module(m)
      return m

    _load("taichi_splatting.torch_lib.transforms", REF + "/torch_lib/transforms.py")
    proj = _load("taichi_splatting.torch_lib.projection", REF + "/torch_lib/projection.py")
    _load("taichi_splatting.torch_lib.rsh", REF + "/torch_lib/rsh.py")
    sh = _load("taichi_splatting.torch_lib.spherical_harmonics", REF + "/torch_lib/spherical_harmonics.py")
    return proj, sh
  finally:
    for k in [k for k in sys.modules if _ours(k)]:
      del sys.modules[k]
    sys.modules.update(saved)
