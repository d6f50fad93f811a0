#!/bin/bash
# ncu evidence for one round (run under gpurun, one GPU): launch list of the bench command + one --set full capture
# per kernel variant.  Outputs land in gpurun_out/; summaries are made here with tools/ncu_summary.py.
set -x
TAG=${1:-r1_v10}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
# headline launch: 512 modes, n=265 -> two-warp variant k_evolve_h<9>
ncu --set full --clock-control none --import-source on -k regex:k_evolve -c 1 -f -o gpurun_out/${TAG}_k_evolve_h9 \
    python tools/run_once.py 1 512 > gpurun_out/${TAG}_h9.log 2>&1
# throughput launch: 4096 modes, n=265 -> k_evolve<9,2,4>
ncu --set full --clock-control none --import-source on -k regex:k_evolve -c 1 -f -o gpurun_out/${TAG}_k_evolve9x4 \
    python tools/run_once.py 1 4096 > gpurun_out/${TAG}_9x4.log 2>&1
# tangent launch: 128 modes x 1 direction, n=265 -> k_evolve_t<9>
ncu --set full --clock-control none --import-source on -k regex:k_evolve_t -c 1 -f -o gpurun_out/${TAG}_k_evolve_t9 \
    python tools/run_fisher_once.py 128 1 > gpurun_out/${TAG}_t9.log 2>&1
ls -la gpurun_out
# the reports (with imported source) exceed what gpurun copies back: extract the csv pages here, drop the reports
for r in k_evolve_h9 k_evolve9x4 k_evolve_t9; do
  ncu -i gpurun_out/${TAG}_${r}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${r}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_${r}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_${r}_source.csv 2>/dev/null
  gzip -f gpurun_out/${TAG}_${r}_source.csv
  rm -f gpurun_out/${TAG}_${r}.ncu-rep
done
du -sh gpurun_out
