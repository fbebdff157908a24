#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/r02/densify_margin.py 2>&1 | tail -12 | tee gpurun_out/r02ae_densify_margin.txt
