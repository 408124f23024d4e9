#!/bin/bash
# round 2, GPU call C: parity suite (new tests), wide-BVH A/B on C4, ncu of the wide traversal kernels
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2c_pytest.log 2>&1
tail -12 gpurun_out/r2c_pytest.log
for w in 0 1; do
  B200PG_WIDE=$w timeout 400 python bench.py --workload mesh_10m --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_c4_wide$w.json 2> gpurun_out/r2c_c4_wide$w.err
done
B200PG_WIDE=1 B200PG_TRACE_SPEC=7 timeout 400 python bench.py --workload mesh_10m --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_c4_wide1_spec7.json 2> gpurun_out/r2c_c4_wide1_spec7.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-28s value %7.1f e2e %7.1f ms/step %6.3f | one-lane step %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f | nodes/ray %.1f" % (f[11:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3, d["roofline"]["traversal"]["per_ray"]["nodes"]))
    except Exception as e: print(f, "failed", e)
PY
B200PG_LANES=1 B200PG_OVERLAP_SHADOW=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shadow' -c 6 -o gpurun_out/r2c_prof_c4_wide python tools/profile_mesh.py 1 > gpurun_out/r2c_ncu_c4.log 2>&1
tail -2 gpurun_out/r2c_ncu_c4.log
B200PG_LANES=1 B200PG_OVERLAP_SHADOW=0 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c_launches_c4.csv python tools/profile_mesh.py 1 > gpurun_out/r2c_ncu_c4_list.log 2>&1
