#!/bin/bash
# r02x: paired panel at 6 CTAs per SM (acc stride 11): parity suite, A/B against HEAD and against 5 CTAs per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest.log
tail -4 gpurun_out/r02x_pytest.log
for i in 1 2; do
  GS_BUILD_VARIANT=_head timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02x_ab.txt
  GS_BUILD_VARIANT=_mb5 timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02x_ab.txt
  timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02x_ab.txt
done
timeout 600 python profiles/bwd_variants.py 20 2>&1 | tail -6 | tee gpurun_out/r02x_bwd_variants.txt
