"""C3: heterogeneous medium. Sample-by-sample parity of the volumetric path vs the oracle on a small grid, then timing at 256^3 / 1024^2."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from conftest import load_package
b = load_package()
from b200pg import api
from oracle_lib import Oracle, develop

orc = Oracle()
for phase in ("hg", "isotropic"):
    sb = b.scenes.cornell_medium(128, 128, spp=8, res=48, phase=phase)
    osc = orc.scene(sb)
    p = api.default_params(); p.max_depth = 8; p.volumetric = 1
    it = api.Integrator(api.Scene.from_builder(sb), p)
    rng = np.random.RandomState(3)
    P = (rng.rand(200000, 3) * 1.6 - 0.8 + [0, 0.8, 0]).astype(np.float32)
    go, gg = osc.grid_lookup(0, P), it.k_grid_lookup(0, P)
    print(phase, "grid lookup max abs diff", float(np.abs(go - gg).max()), "nonzero frac", float((go > 0).mean()))
    pix = rng.randint(0, 128 * 128, 100000).astype(np.uint32); smp = rng.randint(0, 64, 100000).astype(np.uint32)
    Lo = osc.radiance(p, pix, smp); Lg = it.k_radiance(pix, smp)
    err = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-3)
    print(phase, "vol radiance: frac rel err > 1e-3:", float((err > 1e-3).mean()), "means", Lo.mean(0), Lg.mean(0))
    for md in (-1, 2, 3):
        p2 = api.default_params(); p2.max_depth = md; p2.volumetric = 1; p2.rr_depth = 3
        it2 = api.Integrator(api.Scene.from_builder(sb), p2)
        Lo = osc.radiance(p2, pix[:20000], smp[:20000]); Lg = it2.k_radiance(pix[:20000], smp[:20000])
        err = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-3)
        print(phase, "maxDepth", md, "frac rel err > 1e-3:", float((err > 1e-3).mean()), "means", float(Lo.mean()), float(Lg.mean()))
    so, sg = None, None
    it.progression(0, 8); st = it.stats()
    f, sto = osc.render(p, 0, 8)
    print(phase, "stats gpu", {k: st[k] for k in ("paths", "normal_rays", "shadow_rays", "path_length_sum")})
    print(phase, "stats orc", {k: sto[k] for k in ("paths", "normal_rays", "shadow_rays", "path_length_sum")})
    io, ig = develop(f), it.develop()
    print(phase, "image mean", io.mean(), ig.mean(), "relMSE", float(np.mean((io - ig) ** 2 / (io ** 2 + 1e-3))))
# full size timing
t0 = time.time(); sb = b.scenes.cornell_medium(1024, 1024, spp=64, res=256); print("scene build s", time.time() - t0)
p = api.default_params(); p.max_depth = 8; p.volumetric = 1
it = api.Integrator(api.Scene.from_builder(sb), p)
for k in range(3):
    s0 = it.stats(); it.progression(k * 4, 4); s1 = it.stats()
    dt = s1["seconds_total"] - s0["seconds_total"]
    print("C3 step", k, "device s", dt, "Mpaths/s", (s1["paths"] - s0["paths"]) / dt / 1e6, "Mrays/s", (s1["normal_rays"] + s1["shadow_rays"] - s0["normal_rays"] - s0["shadow_rays"]) / dt / 1e6)
print(it.stage_times())
img = it.develop(); print("image mean", img.mean(), "finite", np.isfinite(img).all())
