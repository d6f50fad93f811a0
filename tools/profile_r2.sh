#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): launch list of the bench command, --set full captures of the
# headline kernel (k_evolve_duo, 512 modes), the throughput kernel (k_evolve_lane, 4096 modes) and the table producer.
set -x
TAG=${1:-r2_final}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_evolve_duo --launch-skip 1 -c 1 -f -o gpurun_out/${TAG}_team \
    python tools/run_once.py 3 512 > gpurun_out/${TAG}_team.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_evolve_lane -c 1 -f -o gpurun_out/${TAG}_lane \
    python tools/run_once.py 1 4096 > gpurun_out/${TAG}_lane.log 2>&1
ncu --set full --clock-control none -k regex:k_background -c 1 -f -o gpurun_out/${TAG}_background \
    python tools/time_background.py 128 > gpurun_out/${TAG}_background.log 2>&1
for r in team lane background; do
  ncu -i gpurun_out/${TAG}_${r}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${r}_raw.csv 2>/dev/null
  if [ $r != background ]; then
    ncu -i gpurun_out/${TAG}_${r}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_${r}_source.csv 2>/dev/null
    gzip -f gpurun_out/${TAG}_${r}_source.csv
  fi
  rm -f gpurun_out/${TAG}_${r}.ncu-rep
done
du -sh gpurun_out
