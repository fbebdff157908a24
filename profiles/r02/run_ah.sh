#!/bin/bash
# r02ah: tile count on the auxiliary stream beside the depth sort (GS_COUNT_BESIDE_SORT=1, default) against the count
# behind the sort on the caller's stream (=0): parity suite, A/B, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r02ah_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ah_pytest.log
tail -4 gpurun_out/r02ah_pytest.log
for i in 1 2; do
  GS_COUNT_BESIDE_SORT=0 timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02ah_ab.txt
  GS_COUNT_BESIDE_SORT=1 timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02ah_ab.txt
done
GS_MIN_US=0 timeout 300 python profiles/timeline.py 2>&1 | tail -58 | head -36 | tee gpurun_out/r02ah_timeline.txt
