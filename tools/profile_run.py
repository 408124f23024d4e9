"""Short run for ncu: a few steps of the bench workload (same kernels as bench.py, no oracle).
usage: profile_run.py [caustic|cornell] [steps] [guided|plain]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge
pkg = ge.load_package()
from b200pg import api
name = sys.argv[1] if len(sys.argv) > 1 else "caustic"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
guided = len(sys.argv) > 3 and sys.argv[3] == "guided"
sb = {"caustic": lambda: pkg.scenes.cornell_caustic(1024, 1024), "cornell": lambda: pkg.scenes.cornell_box(512, 512),
      "medium": lambda: pkg.scenes.cornell_medium(1024, 1024, res=256), "mesh": lambda: pkg.scenes.mesh_scene(2048, 2048)}[name]()
scene = api.Scene.from_builder(sb)
p = api.default_params(); p.max_depth = 8
p.volumetric = 1 if name == "medium" else 0
if len(sys.argv) > 4 and sys.argv[4] == "gdist":
    p.guided_distance = 1
if guided:
    p.guiding = 1; p.guide_max_components = 16; p.guide_max_cell_samples = 32768
it = api.Integrator(scene, p)
for k in range(steps):
    if guided:
        it.guiding_mode(True, k > 0)
    it.progression(4 * k, 4)
    if guided:
        t0 = time.time(); n, c = it.train_fused(4); print("train", k, n, c, "wall ms", 1e3 * (time.time() - t0))
st = it.stats()
print("paths", st["paths"], "launches", st["kernel_launches"], "device s", st["seconds_total"], it.stage_times())
