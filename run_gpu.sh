mkdir -p gpurun_out
for tool in racecheck synccheck initcheck; do
  timeout 50 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/$tool.log python tools/sanitize.py 2>&1 | tail -1; echo "$tool rc=${PIPESTATUS[0]}"
  tail -2 gpurun_out/$tool.log
done
