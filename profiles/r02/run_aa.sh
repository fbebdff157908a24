#!/bin/bash
# r02aa: per-warp accumulator rows with 64-splat batches (6 CTAs per SM): parity suite, A/B against the atomics form
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r02aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02aa_pytest.log
tail -4 gpurun_out/r02aa_pytest.log
for i in 1 2; do
  GS_BUILD_VARIANT=_head timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02aa_ab.txt
  timeout 300 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02aa_ab.txt
done
timeout 600 python profiles/bwd_variants.py 20 2>&1 | tail -6 | tee gpurun_out/r02aa_bwd_variants.txt
