cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_workload.py > gpurun_out/r02y_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02y_$tool.log | tail -1) $(grep -c sanitize-ok gpurun_out/r02y_$tool.log)"
done | tee gpurun_out/r02y_sanitizer.txt
