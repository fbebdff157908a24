"""Speed bar for the ordering stage: the reference's own CUDA code (cuda_lib.radix_sort_pairs, its one 48-bit sort of
K (tile | depth) keys -- mapper/tile_mapper.py:148-157 -- compiled for sm_100a into oracle/_ref by oracle/build_ref.py)
against this library's orderings on the same keys at the bench workload's K."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import build_ref
from taichi_splatting_b200 import _lib

ref = build_ref.load_module()
assert ref is not None, "build oracle/_ref first: python oracle/build_ref.py (needs /root/reference)"
dev = torch.device("cuda:0")
k, tiles_n = 3_838_201, 16_384
torch.manual_seed(0)
tiles = torch.randint(0, tiles_n, (k,), dtype=torch.int64, device=dev)
depth_bits = torch.rand(k, device=dev).view(torch.int32).to(torch.int64)
keys = (tiles << 32) | depth_bits
values = torch.arange(k, dtype=torch.int32, device=dev)


def timed(fn, n=20):
  for _ in range(3):
    fn()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / n


t_ref = timed(lambda: ref.radix_sort_pairs(keys, values, 0, 48))
nbytes = _lib.c_size_t()
_lib.call("gs_sort_pairs_workspace_bytes", k, 8, nbytes)
ws = _lib.workspace(nbytes.value, dev)
ko, vo = torch.empty_like(keys), torch.empty_like(values)
stream = _lib.stream_ptr(dev)
t_same = timed(lambda: _lib.call("gs_sort_pairs", _lib.ptr(keys), _lib.ptr(values), _lib.ptr(ko), _lib.ptr(vo), k, 8, 0, 46,
                                 ws.data_ptr(), ws.numel(), stream))
tk = tiles.to(torch.int32)
tko = torch.empty_like(tk)
_lib.call("gs_sort_pairs_workspace_bytes", k, 4, nbytes)
ws4 = _lib.workspace(nbytes.value, dev)
t_tile = timed(lambda: _lib.call("gs_sort_pairs", _lib.ptr(tk), _lib.ptr(values), _lib.ptr(tko), _lib.ptr(vo), k, 4, 0, 14,
                                 ws4.data_ptr(), ws4.numel(), stream))
print(f"K = {k} pairs: reference cuda_lib.radix_sort_pairs(end_bit=48) {t_ref:.3f} ms | gs_sort_pairs same 64-bit keys, populated "
      f"46 bits {t_same:.3f} ms | two-level ordering's tile pass (4-byte keys, 14 bits) {t_tile:.3f} ms (+ 0.062 ms depth order on V)")
