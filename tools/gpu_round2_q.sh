#!/bin/bash
# Round 2, visit Q: compute-sanitizer on the full tour (memcheck, initcheck, synccheck) with the round-2 additions
mkdir -p gpurun_out
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py ) > gpurun_out/memcheck_full.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck_full.log
( time timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize.py ) > gpurun_out/initcheck_full.log 2>&1; echo "initcheck rc=$?"; tail -5 gpurun_out/initcheck_full.log
( time timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize.py ) > gpurun_out/synccheck_full.log 2>&1; echo "synccheck rc=$?"; tail -5 gpurun_out/synccheck_full.log
