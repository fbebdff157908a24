#!/bin/bash
# r02k: where the view-parallel exchange goes at N = 8 (rank-0 GPU timeline), and the e2e arm at N = 8 after the host run-ahead
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 profiles/r02/timeline_multi.py > gpurun_out/r02k_timeline_n8.txt 2> gpurun_out/r02k_timeline_n8.err
cat gpurun_out/r02k_timeline_n8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r02k_cfg3_n8.json 2> gpurun_out/r02k_cfg3_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02k_cfg3_n8.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
PY
