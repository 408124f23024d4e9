"""Ad-hoc GPU-vs-oracle comparison (development aid; the judged tests live in tests/)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from conftest import load_package
b = load_package()
from b200pg import api
from oracle_lib import Oracle, develop

orc = Oracle()
res = int(os.environ.get("RES", "256"))
sb = b.scenes.cornell_box(res, res, spp=16)
osc = orc.scene(sb)
sc = api.Scene.from_builder(sb)
p = api.default_params(); p.max_depth = 8
it = api.Integrator(sc, p)
rng = np.random.RandomState(1)
# --- trace parity
pos = rng.rand(200000, 2).astype(np.float32) * res
rays = osc.camera_rays(pos)
tuv_o, prim_o, cnt = osc.trace(rays)
tuv_g, prim_g = it.k_trace(rays)
hit = prim_o != 0xFFFFFFFF
print("trace: prim mismatch", int((prim_o != prim_g).sum()), "of", len(prim_o), "hit frac", hit.mean())
m = hit & (prim_o == prim_g)
print("  max rel dt", float(np.max(np.abs(tuv_o[m, 0] - tuv_g[m, 0]) / tuv_o[m, 0])), "max duv", float(np.abs(tuv_o[m, 1:] - tuv_g[m, 1:]).max()))
P = rays[m, :3] + rays[m, 4:7] * tuv_o[m, 0:1]
d = rng.randn(P.shape[0], 3).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
r2 = np.concatenate([P, np.full((P.shape[0], 1), 1e-4, np.float32), d, np.full((P.shape[0], 1), np.inf, np.float32)], 1).astype(np.float32)
tuv_o2, prim_o2, _ = osc.trace(r2)
tuv_g2, prim_g2 = it.k_trace(r2)
print("trace2: prim mismatch", int((prim_o2 != prim_g2).sum()), "of", len(prim_o2))
_, sp_o, _ = osc.trace(r2, shadow=True)
_, sp_g = it.k_trace(r2, shadow=True)
print("shadow: mismatch", int(((sp_o != 0xFFFFFFFF) != (sp_g != 0xFFFFFFFF)).sum()))
# --- radiance parity
n = 100000
pix = rng.randint(0, res * res, n).astype(np.uint32); smp = rng.randint(0, 64, n).astype(np.uint32)
Lo = osc.radiance(p, pix, smp)
Lg = it.k_radiance(pix, smp)
err = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-3)
print("radiance: frac rel err > 1e-3:", float((err > 1e-3).mean()), "mean", float(Lo.mean()), float(Lg.mean()))
# --- render
t0 = time.time(); it.progression(0, 16); t_g = time.time() - t0
st = it.stats(); print("gpu stats", st)
film_g = it.film()
film_o, st_o = osc.render(p, 0, 16)
print("oracle stats", st_o)
ig, io = film_g[..., :3] / np.maximum(film_g[..., 4:5], 1e-20), develop(film_o)
rel = np.abs(ig - io).mean() / io.mean()
print("image rel L1", float(rel), "weight max diff", float(np.abs(film_g[..., 4] - film_o[..., 4]).max()))
print("GPU Mpaths/s", st["paths"] / t_g / 1e6, "CPU Mpaths/s", st_o["paths"] / st_o["seconds"] / 1e6)
mm = np.where(prim_o2 != prim_g2)[0]
om = prim_o2[mm] == 0xFFFFFFFF; gm = prim_g2[mm] == 0xFFFFFFFF
print("trace2 mismatch: oracle miss", int(om.sum()), "gpu miss", int(gm.sum()), "both hit", int((~om & ~gm).sum()))
for i in mm[:8]:
    print(r2[i], tuv_o2[i], prim_o2[i], tuv_g2[i], prim_g2[i], "src prim", prim_o[m][i])
m2 = (prim_o2 != 0xFFFFFFFF) & (prim_o2 == prim_g2)
dt = np.abs(tuv_o2[m2, 0] - tuv_g2[m2, 0]); rel = dt / (1 + tuv_o2[m2, 0])
print("secondary dt: max abs", dt.max(), "q99.9 rel(1+t)", np.quantile(rel, 0.999), "q99.99", np.quantile(rel, 0.9999), "max rel", rel.max())
duv = np.abs(tuv_o2[m2, 1:] - tuv_g2[m2, 1:]).max(1)
print("secondary duv: max", duv.max(), "q99.9", np.quantile(duv, 0.999))
