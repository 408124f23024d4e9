#!/bin/bash
# round 2, GPU call S (8 GPUs): BASELINE config 3 -- the 10 M-triangle mesh at 2048^2 on 8 GPUs: weak line (sample batches per GPU) and
# the strong-scaling extra = the single-GPU step with its rows split into 8 bands (tile partition)
mkdir -p gpurun_out
N=${1:-8}
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload mesh_10m --spp-per-step 1 --steps 8 --warmup 3 --pretrain 8) > gpurun_out/r2s_bench_c4_${N}gpu.json 2> gpurun_out/r2s_bench_c4_${N}gpu.err
tail -4 gpurun_out/r2s_bench_c4_${N}gpu.err
