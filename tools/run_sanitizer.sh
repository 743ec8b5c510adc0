#!/bin/bash
# compute-sanitizer over tools/sanitize.py with all four tools; summaries go to gpurun_out/sanitizer_<tool>.log
# (copied to profiles/ for the record). usage: tools/run_sanitizer.sh   (under gpurun, one GPU)
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/sanitizer_$tool.full.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.full.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run done|exit code|Error|hazard" gpurun_out/sanitizer_$tool.full.log | head -40 > gpurun_out/sanitizer_$tool.log
done
tail -n 3 gpurun_out/sanitizer_*.log | grep -v full
