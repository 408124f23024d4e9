#!/bin/bash
# round 2, GPU call J: runtime visit budget (off for small scenes), budget sweep per queue size on C4, C2 / C3 check, ncu of k_trace_tail,
# C3 launch list
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2j_pytest.log 2>&1
tail -4 gpurun_out/r2j_pytest.log
run() { # name workload env...
  n=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2j_$n.json 2> gpurun_out/r2j_$n.err
}
run c2_main cornell_caustic_1024 A=0
run c4_96_96 mesh_10m A=0
run c4_128_48 mesh_10m B200PG_TAIL_VISITS=128 B200PG_TAIL_VISITS_SMALL=48
run c4_128_32 mesh_10m B200PG_TAIL_VISITS=128 B200PG_TAIL_VISITS_SMALL=32
run c4_160_32 mesh_10m B200PG_TAIL_VISITS=160 B200PG_TAIL_VISITS_SMALL=32
run c4_128_16 mesh_10m B200PG_TAIL_VISITS=128 B200PG_TAIL_VISITS_SMALL=16
run c4_128_64 mesh_10m B200PG_TAIL_VISITS=128 B200PG_TAIL_VISITS_SMALL=64
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-26s value %7.1f e2e %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[15:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2j_launches_c3_guided.csv python tools/profile_step.py medium_1024 8 > gpurun_out/r2j_ncu_c3_list.log 2>&1
tail -1 gpurun_out/r2j_ncu_c3_list.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_trace_tail|k_trace_spec' -c 8 -o gpurun_out/r2j_prof_c4_tail python tools/profile_step.py mesh_10m 8 > gpurun_out/r2j_ncu_c4.log 2>&1
tail -1 gpurun_out/r2j_ncu_c4.log
