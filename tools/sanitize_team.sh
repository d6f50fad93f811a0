#!/bin/bash
# compute-sanitizer passes over the CTA-per-mode kernel (run under gpurun, one GPU)
mkdir -p gpurun_out
export DEB_SANITIZE_PRIMAL_ONLY=1 DEB_VARIANT=team
for tool in memcheck racecheck synccheck; do
  echo "==== $tool" >> gpurun_out/r1_sanitizer_team.txt
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "SUMMARY|hazard|Error|error|status|Barrier" | head -40 >> gpurun_out/r1_sanitizer_team.txt
done
cat gpurun_out/r1_sanitizer_team.txt
