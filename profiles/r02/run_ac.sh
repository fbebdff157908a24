#!/bin/bash
# r02ac: GPU timeline of one step (every kernel, idle gaps) on one GPU
mkdir -p gpurun_out
GS_MIN_US=0 timeout 300 python profiles/timeline.py 2>&1 | tail -60 | tee gpurun_out/r02ac_timeline.txt
