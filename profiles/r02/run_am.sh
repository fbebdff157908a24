#!/bin/bash
# r02am: final build after the single-device guard on the mapped count words: parity suite (default switches), the
# fused-path tests again with GS_MAPPED_COUNTS=0 (copy + synchronise, ordered scan without the mapped word) and with
# GS_COUNT_BESIDE_SORT=0, step time
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs > gpurun_out/r02am_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_pytest.log
tail -2 gpurun_out/r02am_pytest.log
for sw in "GS_MAPPED_COUNTS=0" "GS_COUNT_BESIDE_SORT=0"; do
  r=$(env $sw timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused or render_gaussians or cfg5" 2>&1 | tail -1)
  echo "$sw: $r" | tee -a gpurun_out/r02am_switches.txt
done
timeout 200 python profiles/ab_step.py 200 2>&1 | tail -1 | tee -a gpurun_out/r02am_switches.txt
