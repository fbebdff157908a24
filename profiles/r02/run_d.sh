#!/bin/bash
# r02d: barrier-free backward batches (batch 96, double-buffered accumulators), flat pack kernel, shared count/emit query.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -8 gpurun_out/r02d_pytest.log
rm -f gpurun_out/r02d_stages.txt
for v in "" _noexact _skip; do
  echo "=== variant '$v' (per-stage entry points chained from Python)" >> gpurun_out/r02d_stages.txt
  GS_FUSED_HOST=0 GS_BUILD_VARIANT=$v timeout 300 python profiles/time_stages.py 30 2>&1 | grep -E "median|wall" >> gpurun_out/r02d_stages.txt
  echo "--- fused drivers" >> gpurun_out/r02d_stages.txt
  GS_BUILD_VARIANT=$v timeout 300 python profiles/time_stages.py 30 2>&1 | grep -E "wall" >> gpurun_out/r02d_stages.txt
done
cat gpurun_out/r02d_stages.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
python - <<'PY'
import json
for f in ("r02d_bench",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "bwd", d["roofline"]["kernel_ms"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
