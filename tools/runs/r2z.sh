#!/bin/bash
# round 2, GPU call Y (re-used for Z): binning keys from the shade stage
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2z_pytest.log 2>&1
tail -3 gpurun_out/r2z_pytest.log
run() { n=$1; w=$2; shift 2; env "$@" timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-workloads $EXTRA > gpurun_out/r2z_$n.json 2> gpurun_out/r2z_$n.err; }
run c2_main cornell_caustic_1024 A=0
EXTRA="--guided-distance"
run c3_main medium_1024 A=0
EXTRA="--max-cell-samples 4096"
run c2_bigfield cornell_caustic_1024 A=0
EXTRA=
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-22s value %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %5.3f train %5.2f" % (f[15:], d["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
