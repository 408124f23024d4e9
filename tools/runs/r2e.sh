#!/bin/bash
# round 2, GPU call E: parity suite with the warp-cooperative guiding queries in k_shade, C2 A/B (cooperative vs per-thread lobe
# loops, CTAs per SM), C4 with the prefetch off, ncu of one steady-state step of k_shade
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2e_pytest.log 2>&1
tail -6 gpurun_out/r2e_pytest.log
V=$PWD/mitsuba-path-guiding_b200/_variants
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2e_c2_main.json 2> gpurun_out/r2e_c2_main.err
for v in coop0 b6 b5 b10; do
  B200PG_LIB=$V/libb200pg_$v.so timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2e_c2_$v.json 2> gpurun_out/r2e_c2_$v.err
done
timeout 400 python bench.py --workload mesh_10m --steps 6 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2e_c4_main.json 2> gpurun_out/r2e_c4_main.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-24s value %7.1f e2e %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[11:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
B200PG_LANES=1 B200PG_OVERLAP_SHADOW=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name k_shade --launch-skip 135 --launch-count 9 -o gpurun_out/r2e_prof_shade python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2e_ncu_shade.log 2>&1
tail -2 gpurun_out/r2e_ncu_shade.log
