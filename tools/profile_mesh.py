"""C4 (10 M-triangle mesh) short run for ncu. usage: profile_mesh.py [steps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge
pkg = ge.load_package()
from b200pg import api
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sb = pkg.scenes.mesh_scene(2048, 2048)
p = api.default_params(); p.max_depth = 8
it = api.Integrator(api.Scene.from_builder(sb), p)
for k in range(steps):
    it.progression(k, 1)
st = it.stats()
print("paths", st["paths"], "device s", st["seconds_total"], it.stage_times())
