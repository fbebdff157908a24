#!/bin/bash
# r02al: evidence of the FINAL build (1 GPU): parity suite, launch list, --set full of every hand-written kernel of a
# step, summaries regenerated on the box BEFORE the bench line reads them, default bench line, racecheck.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs > gpurun_out/r02al_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02al_pytest.log
tail -3 gpurun_out/r02al_pytest.log
GS_STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02al.csv python profiles/profile_step.py > gpurun_out/r02al_ncu1.log 2>&1
GS_STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'project_|sh_|tile_|depth_key|raster_|finish_scan|DeviceSelectSweep' -s 12 -c 12 -o gpurun_out/raster_r02al -f python profiles/profile_step.py > gpurun_out/r02al_ncu2.log 2>&1
tail -1 gpurun_out/r02al_ncu2.log
python profiles/summarize.py r02al 3 && cp profiles/r02al_summary.md profiles/traffic.json gpurun_out/
timeout 600 python bench.py > gpurun_out/r02al_bench.json 2> gpurun_out/r02al_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02al_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "bwd", d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["limiter"], d["clocks"], d["cpu_baseline"]["ms_per_step"], d["gpu_launches"])
PY
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/r02al_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02al_sanitizer_racecheck.log
tail -3 gpurun_out/r02al_sanitizer_racecheck.log
