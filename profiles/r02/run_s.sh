#!/bin/bash
# r02s: N = 8 A/B of the finished peer-memory exchange (side-stream pack + peer all-reduce) against the NCCL exchange; N = 4 peer
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900 + n)) bench.py --gpus $n --steps 40 --warmup 5 > gpurun_out/$name.json 2> gpurun_out/$name.err; }
run r02s_n8_peer 8 GS_PEER_EXCHANGE=1
run r02s_n8_nccl 8 GS_PEER_EXCHANGE=0
run r02s_n4_peer 4 GS_PEER_EXCHANGE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 profiles/r02/timeline_multi.py 2> gpurun_out/r02s_err.txt | grep -E "nccl|ncclDev|sh_|project_bwd|barrier|allreduce|Memcpy|raster_bwd|step span" > gpurun_out/r02s_timeline_n8.txt; cat gpurun_out/r02s_timeline_n8.txt
python - <<'PY'
import json
for f in ("r02s_n8_peer","r02s_n8_nccl","r02s_n4_peer"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("multi_gpu_check",{}).get("worst_over_ranks"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
