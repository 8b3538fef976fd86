#!/bin/bash
# compute-sanitizer over every launch path (tools/sanitize_paths.py); logs -> gpurun_out/sanitizer/ (copy to profiles/sanitizer/).
# Usage (GPU box): bash tools/run_sanitizer.sh [per-tool timeout in s]
T=${1:-420}
OUT=gpurun_out/sanitizer
mkdir -p $OUT
for tool in memcheck racecheck initcheck synccheck; do
  extra=""
  [ $tool == memcheck ] && extra="--leak-check full"
  [ $tool == initcheck ] && extra=""
  ( time timeout $T compute-sanitizer --tool $tool $extra --print-limit 30 python tools/sanitize_paths.py ) > $OUT/$tool.log 2>&1
  echo "$tool: exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$tool.log | tail -1)"
done
