#!/bin/bash
# r02v: raster backward with output groups switched off (bound for a panel restructure)
mkdir -p gpurun_out
timeout 600 python profiles/bwd_variants.py 20 2>&1 | tail -8 | tee gpurun_out/r02v_bwd_variants.txt
