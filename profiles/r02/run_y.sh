#!/bin/bash
# r02y: ncu --set full of the raster forward / backward kernels (paired panel, 6 CTAs per SM)
mkdir -p gpurun_out
GS_STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_bwd_t|raster_fwd_bulk' -s 2 -c 2 -o gpurun_out/raster_r02y -f python profiles/profile_step.py > gpurun_out/r02y_ncu.log 2>&1
tail -3 gpurun_out/r02y_ncu.log
