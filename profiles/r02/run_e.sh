#!/bin/bash
# r02e: why is the barrier-free backward slower?  full capture of fwd + bwd of the barrier-free build
mkdir -p gpurun_out
GS_STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 4 -c 4 -o gpurun_out/raster_r02e -f python profiles/profile_step.py > gpurun_out/r02e_ncu.log 2>&1
tail -3 gpurun_out/r02e_ncu.log
