#!/bin/bash
# round 2, GPU call P (N GPUs = $1): bench.py under torchrun with N ranks, equal-time relMSE at 4K on N GPUs (in-library device list)
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2p_gpus_$N.txt
[ "$2" = "noband" ] || (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 16 --warmup 3) > gpurun_out/r2p_bench_${N}gpu.json 2> gpurun_out/r2p_bench_${N}gpu.err
[ "$2" = "noband" ] || tail -2 gpurun_out/r2p_bench_${N}gpu.err
timeout 900 python tools/equal_time.py --scene c2 --size 3840x2160 --budgets 10,30 --ref-spp $((4096*N)) --gpus $N > gpurun_out/r2p_equal_time_c5_4k_${N}gpu$3.jsonl 2> gpurun_out/r2p_equal_time_c5_4k_${N}gpu$3.err
tail -2 gpurun_out/r2p_equal_time_c5_4k_${N}gpu$3.err
