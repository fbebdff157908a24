#!/bin/bash
# r02j: single-pass projection (cub::DeviceSelect with the projection as its load), ranges fused into pack, race fix.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -8 gpurun_out/r02j_pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
GS_STEPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02j.csv python profiles/profile_step.py > gpurun_out/r02j_ncu1.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/r02j_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02j_sanitizer_racecheck.log
tail -4 gpurun_out/r02j_sanitizer_racecheck.log
python - <<'PY'
import json,csv
d=json.loads(open("gpurun_out/r02j_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "bwd", d["roofline"]["kernel_ms"], d["clocks"]["sm_mhz"])
rows=list(csv.DictReader([l for l in open('gpurun_out/launches_r02j.csv') if not l.startswith("==")]))
starts=[i for i,r in enumerate(rows) if "DeviceSelectSweep" in r["Kernel Name"] or "project_cull" in r["Kernel Name"]]
rows=rows[starts[-1]-1:]
tot=0
for r in rows:
    v=float(r["Metric Value"].replace(",","")); tot+=v
    print(f'{v/1e3:8.1f}  {r["Kernel Name"][:90]}')
print(tot/1e3)
PY
