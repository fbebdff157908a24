#!/bin/bash
# r02t: evidence of the FINAL build (1 GPU): smoke, parity suite, default bench line, launch list, --set full of every
# hand-written kernel of a step, compute-sanitizer memcheck + racecheck.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02t_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02t_smoke.log; tail -2 gpurun_out/r02t_smoke.log
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t_pytest.log
tail -6 gpurun_out/r02t_pytest.log
timeout 900 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02t_bench_reference.json 2> gpurun_out/r02t_bench_reference.err
GS_STEPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02t.csv python profiles/profile_step.py > gpurun_out/r02t_ncu1.log 2>&1
GS_STEPS=2 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'project_|sh_|tile_|depth_key|raster_|finish_scan|DeviceSelectSweep' -s 13 -c 13 -o gpurun_out/raster_r02t -f python profiles/profile_step.py > gpurun_out/r02t_ncu2.log 2>&1
tail -2 gpurun_out/r02t_ncu2.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/r02t_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02t_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/r02t_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02t_sanitizer_racecheck.log
tail -3 gpurun_out/r02t_sanitizer_memcheck.log gpurun_out/r02t_sanitizer_racecheck.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02t_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "bwd", d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"]["ms_per_step"], d["gpu_launches"])
r=json.loads(open("gpurun_out/r02t_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["ms_per_step"], r["steps"], r["cpu_baseline"]["cores"])
PY
