"""CPU oracle for the Gaussian-splat render hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this package.  taichi_splatting_b200 never imports it, and has no CPU fallback.

  gs_oracle.c / cbind.py  tile mapper + rasteriser fwd/bwd restated in C from the Taichi sources
                          (no CPU version exists in the reference) -- "parity unpinned" against
                          reference outputs (Taichi not installable), pinned by properties.
  torch_ops.py            projection, SH, ndc depth restated from the reference's torch_lib --
                          PINNED by tests/golden/*.npz generated from the real torch_lib.
  random_data.py          the reference's seeded input generators, restated.
  ref_loader.py           imports the reference's own torch_lib from /root/reference (build
                          container only; never used at GPU-box run time).
  pipeline.py             whole render path fwd+bwd on CPU (parity checker + CPU baseline).
"""
