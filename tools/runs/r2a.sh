#!/bin/bash
# round 2, GPU call A: parity suite, C2/C3/C4 benches (trace kernel A/B), ncu captures of the C4 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_c2.json 2> gpurun_out/r2a_c2.err
for spec in 0 3 7; do
  B200PG_TRACE_SPEC=$spec timeout 400 python bench.py --workload mesh_10m --spp-per-step 1 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_c4_spec$spec.json 2> gpurun_out/r2a_c4_spec$spec.err
done
for spec in 0 3; do
  B200PG_TRACE_SPEC=$spec timeout 300 python bench.py --workload medium_1024 --steps 6 --warmup 3 --no-cpu-baseline --guided-distance > gpurun_out/r2a_c3_spec$spec.json 2> gpurun_out/r2a_c3_spec$spec.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2a_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]; n=d["steps"]
        print("%-34s value %7.1f e2e %7.1f ms/step %6.3f | per step: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[11:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["trace"]*1e3/n, s["shade"]*1e3/n, s["shadow"]*1e3/n, s["film"]*1e3/n, s["train"]*1e3/n))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_trace|k_shadow' -c 9 -o gpurun_out/r2a_prof_c4 python tools/profile_mesh.py 1 > gpurun_out/r2a_ncu_c4.log 2>&1
tail -2 gpurun_out/r2a_ncu_c4.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2a_launches_c4.csv python tools/profile_mesh.py 2 > gpurun_out/r2a_ncu_c4_list.log 2>&1
ls -la gpurun_out | tail -15
