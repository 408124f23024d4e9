#!/bin/bash
# round 2, GPU call T: training samples as 32-byte records (one gather per sample): parity suite, C2 (small / 8-GPU-size field), C3, C4
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2t_pytest.log 2>&1
tail -3 gpurun_out/r2t_pytest.log
run() { # name workload env...
  n=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-workloads $EXTRA > gpurun_out/r2t_$n.json 2> gpurun_out/r2t_$n.err
}
run c2_main cornell_caustic_1024 A=0
run c4_main mesh_10m A=0
EXTRA="--guided-distance"
run c3_main medium_1024 A=0
EXTRA="--max-cell-samples 4096"
run c2_bigfield cornell_caustic_1024 A=0
EXTRA=
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-26s value %7.1f e2e %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f | %s" % (f[15:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3, d["config"]["guiding"][-22:]))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2t_launches_c2.csv python tools/profile_step.py cornell_caustic_1024 12 4 > gpurun_out/r2t_ncu_c2.log 2>&1
tail -1 gpurun_out/r2t_ncu_c2.log
