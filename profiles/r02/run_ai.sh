#!/bin/bash
# r02ai: 2 GPUs, final build -- NCCL parity tests, view-parallel and tile-sharded bench lines with their multi_gpu_check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -x > gpurun_out/r02ai_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ai_pytest_2gpu.log
tail -3 gpurun_out/r02ai_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02ai_bench_n2.json 2> gpurun_out/r02ai_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg4 > gpurun_out/r02ai_bench_cfg4_n2.json 2> gpurun_out/r02ai_bench_cfg4_n2.err; echo "bench cfg4 n2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02ai_bench_n2.json", "gpurun_out/r02ai_bench_cfg4_n2.json"):
  d = json.loads(open(f).read().strip().splitlines()[-1])
  print(f, d["ms_per_step"], d["value"], d["multi_gpu_check"]["worst_over_ranks"], d["multi_gpu_check"]["tolerance"])
PY
