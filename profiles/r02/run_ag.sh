#!/bin/bash
# r02ag: gradient accumulators zeroed beside the raster forward (GS_PREZERO=1, default) against memsets in front of
# the raster backward (GS_PREZERO=0): parity suite, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r02ag_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ag_pytest.log
tail -4 gpurun_out/r02ag_pytest.log
for i in 1 2; do
  GS_PREZERO=0 timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02ag_ab.txt
  GS_PREZERO=1 timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02ag_ab.txt
done
