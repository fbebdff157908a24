#!/bin/bash
# r02h: N2/N3/N4 tests, optimiser step in one launch, B-torch-gpu baseline, bench CLIs, e2e with host run-ahead (N=1, 2).
# Run: gpurun --gpus 2 -- bash profiles/r02/run_h.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -8 gpurun_out/r02h_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python profiles/bench_optim.py 30 > gpurun_out/r02h_optim.txt 2>&1; cat gpurun_out/r02h_optim.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python tests/baseline_torch_gpu.py 10 > gpurun_out/r02h_torch_gpu_baseline.txt 2>&1; cat gpurun_out/r02h_torch_gpu_baseline.txt
( CUDA_VISIBLE_DEVICES=0 timeout 300 python -m taichi_splatting_b200.benchmarks.bench_projection --iters 200 --fixed_camera --margin 0.0 --image_size 2048,2048 --n 1000000
  CUDA_VISIBLE_DEVICES=0 timeout 300 python -m taichi_splatting_b200.benchmarks.bench_sh --iters 100
  CUDA_VISIBLE_DEVICES=0 timeout 300 python -m taichi_splatting_b200.benchmarks.bench_tilemapper --iters 200 --reference_sort
  CUDA_VISIBLE_DEVICES=0 timeout 300 python -m taichi_splatting_b200.benchmarks.bench_rasterizer --iters 25 ) > gpurun_out/r02h_bench_clis.txt 2>&1
tail -30 gpurun_out/r02h_bench_clis.txt
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
python - <<'PY'
import json
for f in ("r02h_bench_n1","r02h_bench_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "bwd", d["roofline"]["kernel_ms"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
