"""C4: procedural mesh scene. Parity on a 200k-triangle version vs the oracle; performance + counters at ~10M triangles."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from conftest import load_package
b = load_package()
from b200pg import api
from oracle_lib import Oracle, develop

orc = Oracle()
n_small = int(os.environ.get("NS", "317"))
sb = b.scenes.mesh_scene(256, 256, n=n_small)
t0 = time.time(); osc = orc.scene(sb); print("oracle kd build s", time.time() - t0, osc.kd_info())
t0 = time.time(); sc = api.Scene.from_builder(sb); print("bvh build s", time.time() - t0)
p = api.default_params(); p.max_depth = 8
it = api.Integrator(sc, p)
rng = np.random.RandomState(1)
pos = rng.rand(200000, 2).astype(np.float32) * 256
rays = osc.camera_rays(pos)
tuv_o, prim_o, cnt = osc.trace(rays); tuv_g, prim_g = it.k_trace(rays)
hit = prim_o != 0xFFFFFFFF
m = hit & (prim_o == prim_g)
print("mesh trace: mismatch", int((prim_o != prim_g).sum()), "hit frac", hit.mean(), "q99.9 rel dt", float(np.quantile(np.abs(tuv_o[m,0]-tuv_g[m,0])/(1+tuv_o[m,0]), 0.999)))
pix = rng.randint(0, 256 * 256, 50000).astype(np.uint32); smp = rng.randint(0, 64, 50000).astype(np.uint32)
Lo = osc.radiance(p, pix, smp); Lg = it.k_radiance(pix, smp)
err = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-3)
print("mesh radiance: frac rel err > 1e-3:", float((err > 1e-3).mean()), "means", float(Lo.mean()), float(Lg.mean()))
del it, sc, osc
# ---- full size
n_big = int(os.environ.get("NB", "3163"))
t0 = time.time(); sbig = b.scenes.mesh_scene(2048, 2048, n=n_big); print("python scene s", time.time() - t0)
t0 = time.time(); scb = api.Scene.from_builder(sbig); print("BVH build (10M) s", time.time() - t0)
itb = api.Integrator(scb, p)
for k in range(3):
    s0 = itb.stats(); t0 = time.time(); itb.progression(k, 1); dt = time.time() - t0; s1 = itb.stats()
    print("big step", k, "wall", dt, "device", s1["seconds_total"] - s0["seconds_total"], "Mpaths/s", (s1["paths"] - s0["paths"]) / (s1["seconds_total"] - s0["seconds_total"]) / 1e6,
          "Mrays/s", (s1["normal_rays"] + s1["shadow_rays"] - s0["normal_rays"] - s0["shadow_rays"]) / (s1["seconds_total"] - s0["seconds_total"]) / 1e6)
print(itb.stage_times())
itb.set_option("count_traversal", 1); c0 = itb.stats(); itb.progression(10, 1); c1 = itb.stats()
nr = c1["normal_rays"] - c0["normal_rays"] + c1["shadow_rays"] - c0["shadow_rays"]
print("per ray: nodes", (c1["bvh_nodes_visited"] - c0["bvh_nodes_visited"]) / nr, "prims", (c1["prims_tested"] - c0["prims_tested"]) / nr)
img = itb.develop(); print("image mean", img.mean(), "finite", np.isfinite(img).all())
