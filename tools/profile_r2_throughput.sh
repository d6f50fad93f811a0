#!/bin/bash
# Round-2 evidence for the throughput regime (one GPU): ncu --set full + source page of the one-warp kernel
# k_evolve<9,2,4> on a 4096-mode launch (config 3 shape), and the untimed-by-ncu kernel time of the same launch.
set -x
TAG=${1:-r2_v24}
NK=${2:-4096}
mkdir -p gpurun_out
python tools/run_once.py 3 $NK > gpurun_out/${TAG}_run_once_${NK}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_evolve -c 1 -f -o gpurun_out/${TAG}_k_evolve9x4 \
    python tools/run_once.py 1 $NK > gpurun_out/${TAG}_9x4.log 2>&1
ncu -i gpurun_out/${TAG}_k_evolve9x4.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_evolve9x4_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k_evolve9x4.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_k_evolve9x4_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_k_evolve9x4_source.csv
rm -f gpurun_out/${TAG}_k_evolve9x4.ncu-rep
du -sh gpurun_out
cat gpurun_out/${TAG}_run_once_${NK}.log
