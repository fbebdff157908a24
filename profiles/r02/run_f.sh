#!/bin/bash
# r02f: reverted barrier kernels + tile-sharded cfg4.  Run: gpurun --gpus 2 -- bash profiles/r02/run_f.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -8 gpurun_out/r02f_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02f_bench_cfg3_n1.json 2> gpurun_out/r02f_bench_cfg3_n1.err
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --workload cfg4 --steps 20 --warmup 3 > gpurun_out/r02f_bench_cfg4_n1.json 2> gpurun_out/r02f_bench_cfg4_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload cfg4 --steps 20 --warmup 3 > gpurun_out/r02f_bench_cfg4_n2.json 2> gpurun_out/r02f_bench_cfg4_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02f_bench_cfg3_n2.json 2> gpurun_out/r02f_bench_cfg3_n2.err
python - <<'PY'
import json
for f in ("r02f_bench_cfg3_n1","r02f_bench_cfg4_n1","r02f_bench_cfg4_n2","r02f_bench_cfg3_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "bwd", d["roofline"]["kernel_ms"], d.get("multi_gpu_check",{}).get("worst_over_ranks"), d.get("tile_shards"), d.get("cpu_baseline",{}).get("ms_per_step"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
