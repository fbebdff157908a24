#!/bin/bash
# r02i: evidence of the final build: parity suite, bench line, launch list, --set full of every hand-written kernel of a step,
# compute-sanitizer memcheck + racecheck on the small case, optimiser bench.   Run: gpurun -- bash profiles/r02/run_i.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -6 gpurun_out/r02i_pytest.log
timeout 300 python profiles/bench_optim.py 30 > gpurun_out/r02i_optim.txt 2>&1; head -3 gpurun_out/r02i_optim.txt
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
GS_STEPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02i.csv python profiles/profile_step.py > gpurun_out/r02i_ncu1.log 2>&1
GS_STEPS=2 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'project_|sh_|tile_|depth_key|raster_|finish_scan' -s 14 -c 14 -o gpurun_out/raster_r02i -f python profiles/profile_step.py > gpurun_out/r02i_ncu2.log 2>&1
tail -2 gpurun_out/r02i_ncu2.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/r02i_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02i_sanitizer_memcheck.log
tail -4 gpurun_out/r02i_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/r02i_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02i_sanitizer_racecheck.log
tail -4 gpurun_out/r02i_sanitizer_racecheck.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02i_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"], "bwd", d["roofline"]["kernel_ms"], d["clocks"], d["cpu_baseline"]["ms_per_step"])
PY
