import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from conftest import load_package
b = load_package()
from b200pg import api
from oracle_lib import Oracle
orc = Oracle()
sb = b.scenes.cornell_medium(64, 64, spp=4, res=24, scale_=10.0)
sb.media[0]["method"] = b._abi.MEDIUM_WOODCOCK if os.environ.get("WOOD") else b._abi.MEDIUM_SIMPSON
osc = orc.scene(sb)
rng = np.random.RandomState(7)
pix = rng.randint(0, 64 * 64, 30000).astype(np.uint32); smp = rng.randint(0, 500, 30000).astype(np.uint32)
for md, rr in ((4, 5), (5, 5), (6, 5), (8, 5), (8, 100), (6, 100), (-1, 100)):
    for nee in (1,):
        p = api.default_params(); p.max_depth = md; p.volumetric = 1; p.use_nee = nee; p.rr_depth = rr
        it = api.Integrator(api.Scene.from_builder(sb), p)
        want = osc.radiance(p, pix, smp); got = it.k_radiance(pix, smp)
        err = np.abs(got - want).max(1) / (np.abs(want).max(1) + 1e-3)
        print("maxDepth", md, "rr", rr, "nee", nee, "frac>2e-3", float((err > 2e-3).mean()), "frac>1e-1", float((err > 1e-1).mean()), "means", float(want.mean()), float(got.mean()), "q50 err", float(np.median(err)))
