#!/bin/bash
mkdir -p gpurun_out/r02
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool python profiles/sanitize_patchify.py > gpurun_out/r02/sanitize_patchify_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r02/sanitize_patchify_$tool.txt
done
