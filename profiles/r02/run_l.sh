#!/bin/bash
# r02l: NCCL protocol / algorithm choice for the two collectives of the view-parallel backward at N = 8
mkdir -p gpurun_out
: > gpurun_out/r02l_nccl_variants.txt
for v in "NCCL_PROTO=Simple" "NCCL_PROTO=LL128" "NCCL_ALGO=NVLS" "NCCL_ALGO=NVLS NCCL_PROTO=Simple"; do
  echo "=== $v" >> gpurun_out/r02l_nccl_variants.txt
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 profiles/r02/timeline_multi.py 2> gpurun_out/r02l_err.txt | grep -E "nccl|ncclDev|sh_bwd_views|project_bwd|step span" >> gpurun_out/r02l_nccl_variants.txt
  tail -3 gpurun_out/r02l_err.txt | grep -i "error\|invalid" >> gpurun_out/r02l_nccl_variants.txt
done
cat gpurun_out/r02l_nccl_variants.txt
