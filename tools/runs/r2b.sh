#!/bin/bash
# round 2, GPU call B: parity suite, lanes / shadow-overlap A/B on C2, C3, C4, full default bench (with the workloads block)
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=6) > gpurun_out/r2b_pytest.log 2>&1
tail -8 gpurun_out/r2b_pytest.log
for l in 1 2 4; do
  timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-workloads --lanes $l > gpurun_out/r2b_c2_l$l.json 2> gpurun_out/r2b_c2_l$l.err
  timeout 400 python bench.py --workload mesh_10m --steps 6 --warmup 3 --no-cpu-baseline --lanes $l > gpurun_out/r2b_c4_l$l.json 2> gpurun_out/r2b_c4_l$l.err
  timeout 300 python bench.py --workload medium_1024 --steps 6 --warmup 3 --no-cpu-baseline --guided-distance --lanes $l > gpurun_out/r2b_c3_l$l.json 2> gpurun_out/r2b_c3_l$l.err
done
B200PG_OVERLAP_SHADOW=0 timeout 400 python bench.py --workload mesh_10m --steps 6 --warmup 3 --no-cpu-baseline --lanes 2 > gpurun_out/r2b_c4_l2_noov.json 2> gpurun_out/r2b_c4_l2_noov.err
B200PG_OVERLAP_SHADOW=0 timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-workloads --lanes 2 > gpurun_out/r2b_c2_l2_noov.json 2> gpurun_out/r2b_c2_l2_noov.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-28s value %7.1f e2e %7.1f ms/step %6.3f | one-lane step %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[11:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
(time timeout 900 python bench.py) > gpurun_out/r2b_bench_full.json 2> gpurun_out/r2b_bench_full.err
tail -3 gpurun_out/r2b_bench_full.err
(time timeout 600 python bench.py --impl reference) > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
tail -3 gpurun_out/r2b_bench_ref.err
