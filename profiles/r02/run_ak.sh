#!/bin/bash
# r02ak: is test_second_device_in_one_process sensitive to GS_COUNT_BESIDE_SORT / GS_MAPPED_COUNTS?  (2 GPUs, 3 runs each)
mkdir -p gpurun_out
for sw in "GS_COUNT_BESIDE_SORT=1" "GS_COUNT_BESIDE_SORT=0" "GS_COUNT_BESIDE_SORT=0 GS_MAPPED_COUNTS=0"; do
  for i in 1 2 3; do
    r=$(env $sw timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k second_device 2>&1 | tail -1)
    echo "$sw run $i: $r" | tee -a gpurun_out/r02ak_second_device.txt
  done
done
