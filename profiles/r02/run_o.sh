#!/bin/bash
# r02o: peer-memory exchange with cached symmetric allocations + symm_mem all-reduce of the geometry buffer (2 GPUs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -rs > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -15 gpurun_out/r02o_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29681 profiles/r02/timeline_multi.py 2> gpurun_out/r02o_err.txt | grep -E "nccl|ncclDev|sh_|project_bwd|barrier|symm|multimem|two_shot|all_reduce|Memcpy|step span" > gpurun_out/r02o_timeline_n2.txt; cat gpurun_out/r02o_timeline_n2.txt; tail -5 gpurun_out/r02o_err.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29682 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02o_bench_n2.json 2> gpurun_out/r02o_bench_n2.err
python - <<'PY'
import json
for f in ("r02o_bench_n2",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d.get("multi_gpu_check",{}).get("worst_over_ranks"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
