#!/bin/bash
# round 2, GPU call G: single-phase k_shade with 256-bit loads (the kept form): parity suite, full default bench (CPU legs, workloads
# block), launch list + ncu --set full of one steady-state GUIDED C4 step (one lane)
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2g_pytest.log 2>&1
tail -4 gpurun_out/r2g_pytest.log
(time timeout 900 python bench.py) > gpurun_out/r2g_bench_full.json 2> gpurun_out/r2g_bench_full.err
tail -3 gpurun_out/r2g_bench_full.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_c4_guided.csv python tools/profile_step.py mesh_10m 8 > gpurun_out/r2g_ncu_c4_list.log 2>&1
tail -1 gpurun_out/r2g_ncu_c4_list.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_shade|k_trace|k_shadow' -c 12 -o gpurun_out/r2g_prof_c4_guided python tools/profile_step.py mesh_10m 8 > gpurun_out/r2g_ncu_c4.log 2>&1
tail -1 gpurun_out/r2g_ncu_c4.log
