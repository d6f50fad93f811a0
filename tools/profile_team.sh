#!/bin/bash
# ncu --set full capture (with source) of the CTA-per-mode kernel on the bench workload (run under gpurun, one GPU).
set -x
TAG=${1:-r1_v12}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_evolve_team -c 1 -f -o gpurun_out/${TAG}_k_evolve_team \
    python tools/run_once.py 1 512 > gpurun_out/${TAG}_team.log 2>&1
ncu -i gpurun_out/${TAG}_k_evolve_team.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_evolve_team_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k_evolve_team.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_k_evolve_team_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_k_evolve_team_source.csv
rm -f gpurun_out/${TAG}_k_evolve_team.ncu-rep
du -sh gpurun_out
