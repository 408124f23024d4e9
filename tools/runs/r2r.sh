#!/bin/bash
# round 2, GPU call R: dynamic work fetch in k_estep, L2 fetch granularity A/B
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2r_pytest.log 2>&1
tail -3 gpurun_out/r2r_pytest.log
run() { # name workload env...
  n=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-workloads $EXTRA > gpurun_out/r2r_$n.json 2> gpurun_out/r2r_$n.err
}
run c2_main cornell_caustic_1024 A=0
run c2_l2f32 cornell_caustic_1024 B200PG_L2_FETCH=32
run c2_l2f128 cornell_caustic_1024 B200PG_L2_FETCH=128
run c4_main mesh_10m A=0
run c4_l2f32 mesh_10m B200PG_L2_FETCH=32
run c4_l2f128 mesh_10m B200PG_L2_FETCH=128
run c3_main medium_1024 A=0
run c3_l2f32 medium_1024 B200PG_L2_FETCH=32
EXTRA="--max-cell-samples 4096"
run c2_bigfield cornell_caustic_1024 A=0
run c2_bigfield_l2f32 cornell_caustic_1024 B200PG_L2_FETCH=32
EXTRA=
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2r_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-26s value %7.1f e2e %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f | %s" % (f[15:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3, d["config"]["guiding"][-22:]))
    except Exception as e: print(f, "failed", e)
PY
