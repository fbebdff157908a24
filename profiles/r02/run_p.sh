#!/bin/bash
# r02p: view-parallel scaling with the fused pack + peer-memory all-gather: N = 8, 4, 2 and the N = 8 timeline
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; }
run r02p_cfg3_n8 8 --steps 30 --warmup 5
run r02p_cfg3_n4 4 --steps 30 --warmup 5
run r02p_cfg3_n2 2 --steps 30 --warmup 5
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r02p_cfg3_n1.json 2> gpurun_out/r02p_cfg3_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 profiles/r02/timeline_multi.py 2> gpurun_out/r02p_err.txt | grep -E "nccl|ncclDev|sh_|project_bwd|barrier|raster_bwd|step span" > gpurun_out/r02p_timeline_n8.txt; cat gpurun_out/r02p_timeline_n8.txt
python - <<'PY'
import json
for f in ("r02p_cfg3_n1","r02p_cfg3_n2","r02p_cfg3_n4","r02p_cfg3_n8"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("multi_gpu_check",{}).get("worst_over_ranks"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
