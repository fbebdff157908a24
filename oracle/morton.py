"""CPU oracle of the Morton ordering (N4).  TEST INFRASTRUCTURE ONLY.

numpy restatement of taichi_splatting/misc/morton_sort.py:13-31 (bit spreading), :41-53 (grid cell), :61-70 (code),
:95-125 (grid_at_resolution, argsort = stable radix argsort of the codes).  The reference's kernels are Taichi and
there is no golden vector upstream: parity unpinned, checked by construction (bit interleaving against a slow loop)."""
import numpy as np


def spread_bits64(x):
  x = x.astype(np.uint64) & np.uint64(0x1fffff)
  x = (x | (x << np.uint64(32))) & np.uint64(0x1f00000000ffff)
  x = (x | (x << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
  x = (x | (x << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
  x = (x | (x << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
  x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
  return x


def morton_codes64(points, resolution, size=2**20):
  p = np.asarray(points, dtype=np.float32)
  lower = p.min(axis=0)
  upper = (lower + np.float32(size * resolution)).astype(np.float32)
  inc = ((upper - lower) / np.float32(size)).astype(np.float32)
  v = ((p - lower) / inc).astype(np.float32)
  cell = np.clip(v, np.float32(0), np.float32(size - 1)).astype(np.uint64)
  return spread_bits64(cell[:, 0]) | (spread_bits64(cell[:, 1]) << np.uint64(1)) | (spread_bits64(cell[:, 2]) << np.uint64(2))


def interleave_slow(cx, cy, cz):
  """Bit-by-bit reference for the spreading trick (tests)."""
  code = 0
  for b in range(21):
    code |= ((cx >> b) & 1) << (3 * b) | ((cy >> b) & 1) << (3 * b + 1) | ((cz >> b) & 1) << (3 * b + 2)
  return code


def argsort(points, resolution):
  return np.argsort(morton_codes64(points, resolution), kind="stable").astype(np.int32)
