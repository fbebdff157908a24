#!/bin/bash
# r02ad: V and K published by the device into mapped pinned words and polled by the host (GS_MAPPED_COUNTS=1, default)
# against cudaMemcpyAsync + cudaStreamSynchronize (GS_MAPPED_COUNTS=0): parity suite, A/B, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r02ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ad_pytest.log
tail -4 gpurun_out/r02ad_pytest.log
for i in 1 2; do
  GS_MAPPED_COUNTS=0 timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02ad_ab.txt
  GS_MAPPED_COUNTS=1 timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02ad_ab.txt
done
GS_MIN_US=0 timeout 300 python profiles/timeline.py 2>&1 | tail -56 | head -34 | tee gpurun_out/r02ad_timeline.txt
