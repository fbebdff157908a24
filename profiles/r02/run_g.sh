#!/bin/bash
# r02g: scaling record on one 8xB200 node: cfg3 view-parallel at N = 4, 8; cfg4 tile-sharded at N = 4, 8.
# Run: gpurun --gpus 8 -- bash profiles/r02/run_g.sh
mkdir -p gpurun_out
run() {  # name N args...
  name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
run r02g_cfg3_n8 8 --steps 30 --warmup 5
run r02g_cfg3_n4 4 --steps 30 --warmup 5
run r02g_cfg4_n8 8 --workload cfg4 --steps 20 --warmup 3
run r02g_cfg4_n4 4 --workload cfg4 --steps 20 --warmup 3
python - <<'PY'
import json
for f in ("r02g_cfg3_n4","r02g_cfg3_n8","r02g_cfg4_n4","r02g_cfg4_n8"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "bwd", d["roofline"]["kernel_ms"], d.get("multi_gpu_check",{}).get("worst_over_ranks"), d.get("tile_shards"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
