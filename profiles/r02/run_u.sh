#!/bin/bash
# r02u: component-major pixel pairs + half-width tail loads in the raster kernels: parity suite, then A/B against the
# previous build (libgsplat_b200_head.so), alternating processes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest.log
tail -4 gpurun_out/r02u_pytest.log
for i in 1 2; do
  GS_BUILD_VARIANT=_head timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02u_ab.txt
  timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02u_ab.txt
done
timeout 300 python profiles/time_stages.py 30 2>&1 | head -30 | tee gpurun_out/r02u_stages.txt
