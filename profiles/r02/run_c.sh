#!/bin/bash
# r02c: barrier-free forward batches, variants (unroll 8 / batch 256 / backward pad skip), view-parallel backward through the C driver.
# Run: gpurun --gpus 2 -- bash profiles/r02/run_c.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -8 gpurun_out/r02c_pytest.log
for v in "" _u8 _b256 _skip; do
  echo "=== variant '$v'" >> gpurun_out/r02c_stages.txt
  CUDA_VISIBLE_DEVICES=0 GS_BUILD_VARIANT=$v timeout 300 python profiles/time_stages.py 30 2>&1 | grep -E "raster|wall" >> gpurun_out/r02c_stages.txt
done
cat gpurun_out/r02c_stages.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
GS_VIEW_PARALLEL_STAGED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02c_bench_n2_staged.json 2> gpurun_out/r02c_bench_n2_staged.err
CUDA_VISIBLE_DEVICES=0 GS_STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_fwd -s 1 -c 1 -o gpurun_out/fwd_r02c -f python profiles/profile_step.py > gpurun_out/r02c_ncu.log 2>&1
python - <<'PY'
import json
for f in ("r02c_bench_n2","r02c_bench_n2_staged"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d.get("multi_gpu_check"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-3000:])
PY
