#!/bin/bash
# ncu --set full + source page of the chain-lane kernel (DEB_VARIANT=lane) on a 4096-mode launch (config 3 shape)
set -x
TAG=${1:-r2_lane}
NK=${2:-4096}
export DEB_VARIANT=lane
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_evolve_lane -c 1 -f -o gpurun_out/${TAG} \
    python tools/run_once.py 1 $NK > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_source.csv
rm -f gpurun_out/${TAG}.ncu-rep
cat gpurun_out/${TAG}.log | tail -3
