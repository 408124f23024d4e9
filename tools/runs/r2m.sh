#!/bin/bash
# round 2, GPU call M (2 GPUs): the multi-GPU tests (skipped on one GPU), bench.py under torchrun with 2 ranks, equal-time at 4K on 2 GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2m_gpus.txt
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -k "multi or device_list or cli") > gpurun_out/r2m_pytest.log 2>&1
tail -5 gpurun_out/r2m_pytest.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3) > gpurun_out/r2m_bench_2gpu.json 2> gpurun_out/r2m_bench_2gpu.err
tail -3 gpurun_out/r2m_bench_2gpu.err
timeout 900 python tools/equal_time.py --scene c2 --size 3840x2160 --budgets 10,30 --ref-spp 4096 --gpus 2 > gpurun_out/r2m_equal_time_c5_4k_2gpu.jsonl 2> gpurun_out/r2m_equal_time_c5_4k_2gpu.err
tail -2 gpurun_out/r2m_equal_time_c5_4k_2gpu.err
