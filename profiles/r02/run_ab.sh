#!/bin/bash
# r02ab: forward per-warp visibility rows (_b), backward flush prefetch (_c), both (default) against neither (_a)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ab_pytest.log
tail -4 gpurun_out/r02ab_pytest.log
for i in 1 2; do
  for v in _a _b _c ""; do
    GS_BUILD_VARIANT=$v timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02ab_ab.txt
  done
done
