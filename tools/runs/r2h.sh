#!/bin/bash
# round 2, GPU call H: hit / miss partition of sparse shade queues (adaptive), lane-major enqueue order with stream priorities,
# wide tree on late bounces only: parity suite, C4 / C2 / C3 A/B
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2h_pytest.log 2>&1
tail -4 gpurun_out/r2h_pytest.log
run() { # name workload env...
  n=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2h_$n.json 2> gpurun_out/r2h_$n.err
}
run c4_part0 mesh_10m B200PG_PARTITION=0
run c4_part1 mesh_10m B200PG_PARTITION=1
run c4_part1_lm2 mesh_10m B200PG_LANE_MAJOR=1
run c4_part1_lm4 mesh_10m B200PG_LANE_MAJOR=1 B200PG_LANES=4
run c4_part1_l4 mesh_10m B200PG_LANES=4
run c4_part1_lm4_wide2 mesh_10m B200PG_LANE_MAJOR=1 B200PG_LANES=4 B200PG_WIDE=1 B200PG_WIDE_MIN_PRIMS=1000000 B200PG_WIDE_FROM=2
run c4_part1_wide2 mesh_10m B200PG_WIDE=1 B200PG_WIDE_MIN_PRIMS=1000000 B200PG_WIDE_FROM=2
run c2_part1 cornell_caustic_1024 B200PG_PARTITION=1
run c2_part2 cornell_caustic_1024 B200PG_PARTITION=2
run c2_lm2 cornell_caustic_1024 B200PG_LANE_MAJOR=1
run c2_lm4 cornell_caustic_1024 B200PG_LANE_MAJOR=1 B200PG_LANES=4
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-26s value %7.1f e2e %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[15:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
