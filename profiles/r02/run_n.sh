#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 profiles/r02/timeline_multi.py > gpurun_out/r02n_timeline_n2.txt 2> gpurun_out/r02n_err.txt; cat gpurun_out/r02n_timeline_n2.txt
