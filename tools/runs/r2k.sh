#!/bin/bash
# round 2, GPU call K: event partition in front of k_shade_vol (C3): parity suite, C3 A/B, C2 / C4 check, full default bench
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2k_pytest.log 2>&1
tail -4 gpurun_out/r2k_pytest.log
run() { # name workload env...
  n=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-workloads $EXTRA > gpurun_out/r2k_$n.json 2> gpurun_out/r2k_$n.err
}
EXTRA=--guided-distance
run c3_part1 medium_1024 A=0
run c3_part0 medium_1024 B200PG_PARTITION=0
EXTRA=
run c3ng_part1 medium_1024 A=0
run c3ng_part0 medium_1024 B200PG_PARTITION=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2k_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-26s value %7.1f e2e %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[15:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
(time timeout 900 python bench.py) > gpurun_out/r2k_bench_full.json 2> gpurun_out/r2k_bench_full.err
tail -3 gpurun_out/r2k_bench_full.err
