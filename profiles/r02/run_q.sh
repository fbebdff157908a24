#!/bin/bash
# r02q: A/B at N = 8 in one call: peer-memory exchange vs NCCL all-gather; e2e with sharded vs full upload
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/$name.json 2> gpurun_out/$name.err; }
run r02q_n8_nccl_full GS_PEER_EXCHANGE=0 GS_E2E_UPLOAD=full
run r02q_n8_peer_sharded GS_PEER_EXCHANGE=1 GS_E2E_UPLOAD=sharded
run r02q_n8_nccl_sharded GS_PEER_EXCHANGE=0 GS_E2E_UPLOAD=sharded
run r02q_n8_peer_full GS_PEER_EXCHANGE=1 GS_E2E_UPLOAD=full
python - <<'PY'
import json
for f in ("r02q_n8_nccl_full","r02q_n8_peer_sharded","r02q_n8_nccl_sharded","r02q_n8_peer_full"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("multi_gpu_check",{}).get("worst_over_ranks"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
