#!/bin/bash
# round 2, GPU call I: visit budget + warp-cooperative tail kernel (k_trace_tail), BSDF single call sites: parity suite, A/B on C4 / C2
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2i_pytest.log 2>&1
tail -4 gpurun_out/r2i_pytest.log
V=$PWD/mitsuba-path-guiding_b200/_variants
run() { # name workload env...
  n=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2i_$n.json 2> gpurun_out/r2i_$n.err
}
run c4_main mesh_10m A=0
run c4_main_spec7 mesh_10m B200PG_TRACE_SPEC=7
for v in tv0 tv64 tv160 tv96_l0 tv96_l12; do run c4_$v mesh_10m B200PG_LIB=$V/libb200pg_$v.so; done
run c4_main_wide2 mesh_10m B200PG_WIDE=1 B200PG_WIDE_MIN_PRIMS=1000000 B200PG_WIDE_FROM=2
run c2_main cornell_caustic_1024 A=0
run c2_tv0 cornell_caustic_1024 B200PG_LIB=$V/libb200pg_tv0.so
run c3_main medium_1024 A=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-26s value %7.1f e2e %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[15:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_c4_guided.csv python tools/profile_step.py mesh_10m 8 > gpurun_out/r2i_ncu_c4_list.log 2>&1
tail -1 gpurun_out/r2i_ncu_c4_list.log
