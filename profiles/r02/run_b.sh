#!/bin/bash
# r02b: bulk-copy staged raster kernels: parity suite + bench + launch list + full capture of the raster kernels.
# Run: gpurun -- bash profiles/r02/run_b.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -15 gpurun_out/r02b_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
GS_RASTER_STAGING=gather timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r02b_bench_gatherfwd.json 2> gpurun_out/r02b_bench_gatherfwd.err
GS_STEPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02b.csv python profiles/profile_step.py > gpurun_out/r02b_ncu1.log 2>&1
GS_STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 3 -c 3 -o gpurun_out/raster_r02b -f python profiles/profile_step.py > gpurun_out/r02b_ncu2.log 2>&1
python - <<'PY'
import json
for f in ("r02b_bench","r02b_bench_gatherfwd"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "bwd", d["roofline"]["kernel_ms"], d["clocks"]["sm_mhz"], {k:v for k,v in d["stages_ms"].items() if "raster" in k})
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
