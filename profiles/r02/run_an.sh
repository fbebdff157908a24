#!/bin/bash
# r02an: compute-sanitizer initcheck (reads of uninitialised device memory) on the small end-to-end case; the caching
# allocator is switched off so that every torch tensor is its own cudaMalloc
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 100 compute-sanitizer --tool initcheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/r02an_initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/r02an_initcheck.log
grep -c "Uninitialized" gpurun_out/r02an_initcheck.log; grep -A3 "Uninitialized" gpurun_out/r02an_initcheck.log | grep -E "Uninitialized| at | in " | sort | uniq -c | sort -rn | head -12; tail -3 gpurun_out/r02an_initcheck.log
