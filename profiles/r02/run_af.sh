#!/bin/bash
# r02af: 2 GPUs -- NCCL parity tests and the view-parallel / tile-sharded bench lines with their multi_gpu_check, on the
# build with the paired backward panel and the mapped V / K words
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -x > gpurun_out/r02af_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02af_pytest_2gpu.log
tail -4 gpurun_out/r02af_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02af_bench_n2.json 2> gpurun_out/r02af_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r02af_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg4 > gpurun_out/r02af_bench_cfg4_n2.json 2> gpurun_out/r02af_bench_cfg4_n2.err; echo "bench cfg4 n2 rc=$?"
tail -c 800 gpurun_out/r02af_bench_cfg4_n2.json
