#!/bin/bash
# round 2, GPU call D: parity suite, prefetch A/B on C4, equal-time sweeps (oracle reference at 512^2 with the CPU arm; 4K), full bench
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2d_pytest.log 2>&1
tail -6 gpurun_out/r2d_pytest.log
timeout 400 python bench.py --workload mesh_10m --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_c4_prefetch.json 2> gpurun_out/r2d_c4_prefetch.err
B200PG_LIB=$PWD/mitsuba-path-guiding_b200/_variants/libb200pg_noprefetch.so timeout 400 python bench.py --workload mesh_10m --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_c4_noprefetch.json 2> gpurun_out/r2d_c4_noprefetch.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_c4*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-28s value %7.1f ms/step %6.3f | one-lane step %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[11:], d["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 python tools/equal_time.py --scene c2 --size 512 --budgets 1,2,4 --cpu > gpurun_out/r2d_equal_time_c2_512.jsonl 2> gpurun_out/r2d_equal_time_c2_512.err
tail -2 gpurun_out/r2d_equal_time_c2_512.err
timeout 900 python tools/equal_time.py --scene c2 --size 3840x2160 --budgets 10,30 --ref-spp 4096 > gpurun_out/r2d_equal_time_c5_4k.jsonl 2> gpurun_out/r2d_equal_time_c5_4k.err
tail -2 gpurun_out/r2d_equal_time_c5_4k.err
(time timeout 900 python bench.py) > gpurun_out/r2d_bench_full.json 2> gpurun_out/r2d_bench_full.err
tail -3 gpurun_out/r2d_bench_full.err
