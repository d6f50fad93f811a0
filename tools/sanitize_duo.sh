#!/bin/bash
# compute-sanitizer passes over the two-team CTA kernel (k_evolve_duo) forced at small sizes (DEB_DUO=2): 6, 4 and 21 modes
# (odd counts: one team of a CTA leaves at once or runs alone), every named-barrier id of both teams in use.
mkdir -p gpurun_out
OUT=gpurun_out/r2_sanitizer_duo.txt
: > $OUT
export DEB_SANITIZE_PRIMAL_ONLY=1 DEB_VARIANT=team DEB_DUO=2
for tool in memcheck synccheck racecheck; do
  echo "==== k_evolve_duo: $tool" >> $OUT
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "SUMMARY|hazard|Error|error|status|Barrier|deb_team" | sort | uniq -c | sort -rn | head -30 >> $OUT
done
cat $OUT
