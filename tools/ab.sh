#!/bin/bash
# A/B of library variants on the GPU box: tools/ab.sh "<bench args>" name1 name2 ... ("default" = the shipped library)
args=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset B200PG_LIB; else export B200PG_LIB=$PWD/mitsuba-path-guiding_b200/_variants/libb200pg_$v.so; fi
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline $args $EXTRA > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/ab_%s.json"%v)); s=d["roofline"]["stage_seconds"]
    print("%-10s value %7.1f e2e %7.1f ms/step %6.3f | trace %5.1f shade %5.1f shadow %4.1f film %4.1f train %5.1f ms/16" % (v, d["value"], d["e2e"]["value"], d["ms_per_step"], s["trace"]*1e3, s["shade"]*1e3, s["shadow"]*1e3, s["film"]*1e3, s["train"]*1e3))
except Exception as e: print(v, "failed", e)
PY
done
