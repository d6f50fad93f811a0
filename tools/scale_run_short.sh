#!/bin/bash
# bench.py and the strong-scaling configs at N GPUs (N = $1; run with gpurun --gpus N)
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_r1_v20_bench_${N}gpu.json 2> gpurun_out/scale_bench_$N.err
tail -c 300 gpurun_out/scale_bench_$N.err; cut -c1-220 gpurun_out/scale_r1_v20_bench_${N}gpu.json
CFG4_MIN_WORLD=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) tools/run_scaling.py > gpurun_out/scale_r1_v20_configs34_${N}gpu.jsonl 2> gpurun_out/scale_cfg_$N.err
tail -c 300 gpurun_out/scale_cfg_$N.err; cat gpurun_out/scale_r1_v20_configs34_${N}gpu.jsonl
