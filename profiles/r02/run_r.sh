#!/bin/bash
# r02r: peer mode with side-stream pack + our own peer all-reduce (2 GPUs: correctness, timeline, bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -rs > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -12 gpurun_out/r02r_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29881 profiles/r02/timeline_multi.py 2> gpurun_out/r02r_err.txt | grep -E "nccl|ncclDev|sh_|project_bwd|barrier|allreduce|Memcpy|raster_bwd|step span" > gpurun_out/r02r_timeline_n2.txt; cat gpurun_out/r02r_timeline_n2.txt; tail -3 gpurun_out/r02r_err.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29882 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r02r_bench_n2.json 2> gpurun_out/r02r_bench_n2.err
GS_PEER_EXCHANGE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29883 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r02r_bench_n2_nccl.json 2> gpurun_out/r02r_bench_n2_nccl.err
python - <<'PY'
import json
for f in ("r02r_bench_n2","r02r_bench_n2_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d.get("multi_gpu_check",{}).get("worst_over_ranks"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
