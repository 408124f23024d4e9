"""Short run for ncu: a few progressions of the bench workload (same kernels as bench.py, no oracle)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge
pkg = ge.load_package()
from b200pg import api
name = sys.argv[1] if len(sys.argv) > 1 else "caustic"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sb = pkg.scenes.cornell_caustic(1024, 1024) if name == "caustic" else pkg.scenes.cornell_box(512, 512)
scene = api.Scene.from_builder(sb)
p = api.default_params(); p.max_depth = 8
it = api.Integrator(scene, p)
for k in range(steps):
    it.progression(4 * k, 4)
st = it.stats()
print("paths", st["paths"], "launches", st["kernel_launches"], "device s", st["seconds_total"], it.stage_times())
