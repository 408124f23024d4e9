#!/bin/bash
# round 2, GPU call N: parity suite (grazing-ray test reworked), lane count on C4, equal-time sweeps with the debiased relMSE
# (512^2 against the oracle fixture with the CPU arm; 4K on one GPU), full bench with pinned scene uploads
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/r2n_pytest.log 2>&1
tail -4 gpurun_out/r2n_pytest.log
run() { # name workload env...
  n=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2n_$n.json 2> gpurun_out/r2n_$n.err
}
run c4_l1 mesh_10m B200PG_LANES=1
run c4_l2 mesh_10m B200PG_LANES=2
run c4_l3 mesh_10m B200PG_LANES=3
run c4_l2_noov mesh_10m B200PG_LANES=2 B200PG_OVERLAP_SHADOW=0
run c2_l1 cornell_caustic_1024 B200PG_LANES=1
run c2_l3 cornell_caustic_1024 B200PG_LANES=3
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n_c*.json")):
    try:
        d=json.load(open(f)); s=d["roofline"]["stage_seconds"]
        print("%-26s value %7.1f e2e %7.1f ms/step %6.3f | one-lane %6.3f: trace %5.2f shade %5.2f shadow %4.2f film %4.2f train %5.2f" % (f[15:], d["value"], d["e2e"]["value"], d["ms_per_step"], s["one_lane_step"]*1e3, s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 python tools/equal_time.py --scene c2 --size 512 --budgets 1,2,4 --cpu > gpurun_out/r2n_equal_time_c2_512.jsonl 2> gpurun_out/r2n_equal_time_c2_512.err
tail -1 gpurun_out/r2n_equal_time_c2_512.err
timeout 900 python tools/equal_time.py --scene c2 --size 3840x2160 --budgets 10,30 --ref-spp 8192 > gpurun_out/r2n_equal_time_c5_4k_1gpu.jsonl 2> gpurun_out/r2n_equal_time_c5_4k_1gpu.err
tail -1 gpurun_out/r2n_equal_time_c5_4k_1gpu.err
(time timeout 900 python bench.py) > gpurun_out/r2n_bench_full.json 2> gpurun_out/r2n_bench_full.err
tail -3 gpurun_out/r2n_bench_full.err
