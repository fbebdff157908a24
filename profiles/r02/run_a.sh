#!/bin/bash
# r02a: parity suite (incl. 2-GPU NCCL tests), bench at eps 0 / 1e-6, bench at N=2.  Run: gpurun --gpus 2 -- bash profiles/r02/run_a.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q -x -rs --durations=15 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -40 gpurun_out/r02a_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r02a_bench_eps0.json 2> gpurun_out/r02a_bench_eps0.err; tail -c 1500 gpurun_out/r02a_bench_eps0.json
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 30 --warmup 5 --fwd-eps 1e-6 > gpurun_out/r02a_bench_eps1e-6.json 2> gpurun_out/r02a_bench_eps1e-6.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02a_bench_n2.json 2> gpurun_out/r02a_bench_n2.err
python - <<'PY'
import json
for f in ("r02a_bench_eps0","r02a_bench_eps1e-6","r02a_bench_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["clocks"], {k:v for k,v in d["stages_ms"].items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
