#!/bin/bash
# ncu evidence for the headline kernel after the two-team CTA change (run under gpurun, one GPU): launch list of the bench
# command and one --set full capture of k_evolve_duo in steady state (second launch: learned work list in use).
TAG=${1:-r2_duo}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_evolve_duo --launch-skip 1 -c 1 -f -o gpurun_out/${TAG}_duo \
    python tools/run_once.py 3 512 > gpurun_out/${TAG}_duo.log 2>&1
ncu -i gpurun_out/${TAG}_duo.ncu-rep --page raw --csv > gpurun_out/${TAG}_duo_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_duo.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_duo_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_duo_source.csv
rm -f gpurun_out/${TAG}_duo.ncu-rep
du -sh gpurun_out
