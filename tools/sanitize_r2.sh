#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (run under gpurun, one GPU): the chain-lane kernel with lock-step
# barriers (more modes than warps per CTA, so warps take new modes and leave the lock-step group at different times), and
# the table producer.
mkdir -p gpurun_out
OUT=gpurun_out/r2_sanitizer.txt
: > $OUT
export DEB_SANITIZE_PRIMAL_ONLY=1
for tool in memcheck racecheck synccheck; do
  echo "==== lane kernel: $tool" >> $OUT
  DEB_VARIANT=lane timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "SUMMARY|hazard|Error|error|status|Barrier" | head -40 >> $OUT
done
for tool in memcheck racecheck; do
  echo "==== background kernel: $tool" >> $OUT
  timeout 900 compute-sanitizer --tool $tool python -c "
import sys; sys.path.insert(0,'disco-eb_b200'); sys.path.insert(0,'tests')
import numpy as np
from discoeb_b200 import _cabi
from discoeb_b200.background import config4_draws, pack_background_input
base = dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0, A_s=2.1064e-09, n_s=0.96822, H0=67.742, Tcmb=2.7255, YHe=0.248, Neff=2.046, Nmnu=1, mnu=0.06)
bg = np.stack([pack_background_input({**base, **d}) for d in config4_draws(3)])
s, t, ms, ex = _cabi.default_library().background_host(bg, 256, extras=True)
print('status ok', np.isfinite(t).all(), np.isfinite(ex[:, 256*7:256*8]).all())
" 2>&1 | grep -E "SUMMARY|hazard|Error|error|status|Barrier" | head -20 >> $OUT
done
cat $OUT
