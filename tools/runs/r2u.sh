#!/bin/bash
# round 2, GPU call U: final state -- build check, smoke(), parity suite, default bench (both arms), launch list of the bench command
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -1 gpurun_out/r2u_smoke.log
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2u_pytest.log 2>&1
tail -3 gpurun_out/r2u_pytest.log
(time timeout 900 python bench.py) > gpurun_out/r2u_bench_full.json 2> gpurun_out/r2u_bench_full.err
tail -3 gpurun_out/r2u_bench_full.err
(time timeout 600 python bench.py --impl reference --steps 4 --warmup 1) > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2u_bench_launches.csv python bench.py --steps 4 --warmup 3 --pretrain 4 --no-cpu-baseline --no-workloads > gpurun_out/r2u_ncu_bench.log 2>&1
tail -1 gpurun_out/r2u_ncu_bench.log | head -c 300
