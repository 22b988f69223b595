#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_case.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok$|Error|error" | head -8
done > gpurun_out/sanitizer_${1:-r2}.txt 2>&1
cat gpurun_out/sanitizer_${1:-r2}.txt
