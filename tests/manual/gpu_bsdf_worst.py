import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from conftest import load_package
b = load_package()
from b200pg import api
from oracle_lib import Oracle
from bsdf_cases import bsdf_scene, random_dirs
orc = Oracle()
sb, idx = bsdf_scene()
osc = orc.scene(sb)
p = api.default_params(); p.max_depth = 8
it = api.Integrator(api.Scene.from_builder(sb), p)
rng = np.random.RandomState(2)
n = 50000
for name, i in idx.items():
    wi = random_dirs(rng, n); wo = random_dirs(rng, n); u = rng.rand(n, 2).astype(np.float32)
    o = osc.bsdf(i, wi, wo, u); g = it.k_bsdf(i, wi, wo, u)
    same = o["flags"] == g["flags"]; ok = same & (o["spdf"] > 0)
    werr = (np.abs(o["weight"] - g["weight"]) / np.maximum(np.abs(o["weight"]), 1e-2)).max(1)
    werr[~ok] = 0
    k = int(np.argmax(werr))
    print(name, "max werr", werr[k], "wi", wi[k], "u", u[k], "\n   o: wo", o["wo"][k], "w", o["weight"][k], "pdf", o["spdf"][k], "fl", hex(o["flags"][k]),
          "\n   g: wo", g["wo"][k], "w", g["weight"][k], "pdf", g["spdf"][k], "fl", hex(g["flags"][k]))
