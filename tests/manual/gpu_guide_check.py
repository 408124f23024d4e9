"""Ad-hoc guiding check: train on the GPU, compare field/query/binning/EM against the oracle, measure relMSE."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from conftest import load_package
b = load_package()
from b200pg import api
from oracle_lib import Oracle, develop

orc = Oracle()
res = int(os.environ.get("RES", "256"))
sb = b.scenes.cornell_caustic(res, res, spp=16)
osc = orc.scene(sb)
sc = api.Scene.from_builder(sb)
p = api.default_params(); p.max_depth = 8; p.guiding = 1; p.guide_max_components = 16; p.guide_max_cell_samples = 20000
it = api.Integrator(sc, p)
# --- train on GPU: 5 progressions x 4 spp
for k in range(5):
    it.guiding_mode(True, k > 0)
    it.progression(100 * k, 4)
    t0 = time.time(); n, c = it.train(4); dt = time.time() - t0
    print("train", k, "samples", n, "cells before", c, "after", it.stats()["guide_cells"], "s", round(dt, 4))
snap = it.field_snapshot()
print("snapshot words", snap.size, "header", snap[:4])
fld = orc.field(16, (0, 0, 0), (1, 1, 1)); fld.load(snap)
rng = np.random.RandomState(0)
n = 200000
pos = (rng.rand(n, 3) * [2, 2, 2] - [1, 0, 1]).astype(np.float32)
d = rng.randn(n, 3).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
u = rng.rand(n, 3).astype(np.float32)
qo = fld.pdf_sample(pos, d, u); qg = it.k_vmm_pdf_sample(pos, d, u)
print("cells equal", np.array_equal(qo["cell"], qg["cell"]), "pdf max rel", float(np.max(np.abs(qo["pdf"] - qg["pdf"]) / np.maximum(qo["pdf"], 1e-3))),
      "dir max abs", float(np.abs(qo["dir"] - qg["dir"]).max()), "spdf max rel", float(np.max(np.abs(qo["spdf"] - qg["spdf"]) / np.maximum(qo["spdf"], 1e-3))))
# binning
co, po, oo = fld.bin(pos); cg, pg_, og = it.k_bin_samples(pos, fld.info()["cells"])
print("bin: cell eq", np.array_equal(co, cg), "perm eq", np.array_equal(po, pg_), "offsets eq", np.array_equal(oo, og))
# guided radiance parity
it.guiding_mode(False, True)
m = 60000
pix = rng.randint(0, res * res, m).astype(np.uint32); smp = rng.randint(0, 64, m).astype(np.uint32)
Lo = osc.radiance(p, pix, smp, field=fld); Lg = it.k_radiance(pix, smp)
err = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-3)
print("guided radiance: frac rel err > 1e-3:", float((err > 1e-3).mean()), "means", float(Lo.mean()), float(Lg.mean()))
# training-sample parity + EM parity on the same samples
sink = orc.samples(); osc.radiance(p, pix, smp, field=fld, sink=sink); s = sink.get()
print("oracle samples", len(s["weight"]))
info = fld.info()
st_o = fld.estep(s); st_g = it.k_em_step(s, 0, info["cells"], info["K"])
den = np.maximum(np.abs(st_o), 1e-3 * np.abs(st_o).max())
print("E-step stats max rel", float((np.abs(st_o - st_g) / den).max()))
fld.train(s, 4, 1e9); it.k_em_step(s, 4, info["cells"], info["K"])
a = fld.snapshot().view(np.float32); g = it.field_snapshot().view(np.float32)
print("after EM: same size", a.size == g.size, "max rel diff lobes", float((np.abs(a[8:] - g[8:]) / np.maximum(np.abs(a[8:]), 1e-3)).max()) if a.size == g.size else None)
# equal-spp relMSE: guided vs unguided
ref, _ = osc.render(p, 5000, 512); ref = develop(ref)
def relmse(img):
    e = ((img - ref) ** 2 / (ref ** 2 + 1e-3)).mean(2).ravel(); e.sort(); return float(e[: int(len(e) * 0.999)].mean())
it.field_load(snap)
it.film_clear(); it.guiding_mode(False, True); t0 = time.time(); it.progression(0, 16); tg = time.time() - t0; img_g = it.develop()
it.film_clear(); it.guiding_mode(False, False); t0 = time.time(); it.progression(0, 16); tu = time.time() - t0; img_u = it.develop()
print("relMSE 16spp guided", relmse(img_g), "unguided", relmse(img_u), "means", float(img_g.mean()), float(img_u.mean()), float(ref.mean()), "times", tg, tu)
for rep in range(3):
    it.film_clear(); it.guiding_mode(False, True); s0 = it.stats(); t0 = time.time(); it.progression(0, 16); tg = time.time() - t0; s1 = it.stats()
    print("guided rep", rep, "wall", tg, "device", s1["seconds_total"] - s0["seconds_total"], it.stage_times())
it.guiding_mode(True, True); t0 = time.time(); it.progression(0, 16); print("guided+record wall", time.time() - t0)
