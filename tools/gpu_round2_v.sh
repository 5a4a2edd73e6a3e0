#!/bin/bash
# Round 2, visit V: compute-sanitizer on the tiny tour with the final kernels (tail kernel, shared-memory backlog, binned queues)
mkdir -p gpurun_out
( time RPTR_CUDA_LIB=variants/librptr_cuda_t128.so timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/sanitize.py --tiny ) > gpurun_out/racecheck_tiny.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/racecheck_tiny.log
( time RPTR_CUDA_LIB=variants/librptr_cuda_t128.so timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py --tiny ) > gpurun_out/memcheck_tiny.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_tiny.log
( time RPTR_CUDA_LIB=variants/librptr_cuda_t128.so timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize.py --tiny ) > gpurun_out/synccheck_tiny.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/synccheck_tiny.log
( time RPTR_CUDA_LIB=variants/librptr_cuda_t128.so timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize.py --tiny ) > gpurun_out/initcheck_tiny.log 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/initcheck_tiny.log
