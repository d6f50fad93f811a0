#!/bin/bash
# bench.py at N = 1, 2, 4, 8 and the strong-scaling configs at N = 8 (and 1 for config 3); run with gpurun --gpus 8
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_bench_$N.json 2> gpurun_out/scale_bench_$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_bench_$N.json 2> gpurun_out/scale_bench_$N.err; fi
  tail -c 300 gpurun_out/scale_bench_$N.err; cut -c1-200 gpurun_out/scale_bench_$N.json
done
for N in 8 2; do
  CFG4_MIN_WORLD=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) tools/run_scaling.py > gpurun_out/scale_cfg_$N.jsonl 2> gpurun_out/scale_cfg_$N.err
  tail -c 300 gpurun_out/scale_cfg_$N.err; cat gpurun_out/scale_cfg_$N.jsonl
done
CFG4_MIN_WORLD=8 python tools/run_scaling.py > gpurun_out/scale_cfg_1.jsonl 2> gpurun_out/scale_cfg_1.err; cat gpurun_out/scale_cfg_1.jsonl
