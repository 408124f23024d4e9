"""relMSE at equal time (north-star metric): guided (training included in the budget) vs unguided, config C2.
usage: equal_time.py [size | WxH] [budgets_s...]   -> one JSON line per budget on stdout"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge
pkg = ge.load_package()
from b200pg import api

size = sys.argv[1] if len(sys.argv) > 1 else "1024"
W, H = (int(x) for x in size.split("x")) if "x" in size else (int(size), int(size))
budgets = [float(x) for x in sys.argv[2:]] or [0.25, 0.5, 1.0, 2.0]
sb = pkg.scenes.cornell_caustic(W, H)
scene = api.Scene.from_builder(sb)


def params(guided):
    p = api.default_params(); p.max_depth = 8
    p.guiding = 1 if guided else 0; p.guide_max_components = 16; p.guide_max_cell_samples = 32768
    return p


def relmse(img, ref):
    e = ((img - ref) ** 2 / (ref ** 2 + 1e-3)).mean(2).ravel()
    e.sort()
    return float(e[: int(len(e) * 0.999)].mean())  # 0.1% outliers trimmed


# converged reference: unguided path tracer, disjoint sample indices (REF_SPP samples per pixel)
it = api.Integrator(scene, params(False))
t0 = time.perf_counter()
ref_spp = int(os.environ.get("REF_SPP", "16384"))
for k in range(ref_spp // 64):
    it.progression(1_000_000 + 64 * k, 64)
ref = it.develop()
print("reference: %d spp in %.1f s" % (ref_spp, time.perf_counter() - t0), file=sys.stderr)
it.close()

spp = 4
for T in budgets:
    out = {"budget_s": T, "size": size}
    for guided in (False, True):
        it = api.Integrator(scene, params(guided))
        it.progression(0, 1); it.film_clear()  # warm the allocations outside the budget
        if guided:
            it.guiding_mode(True, False); it.progression(0, 1); it.train_fused(4); it.film_clear()
            it2 = api.Integrator(scene, params(True)); it.close(); it = it2  # fresh field, warm library
        t0 = time.perf_counter()
        k = 0
        train_until = 0.3 * T  # guided: the first 30% of the budget trains (samples still go to the film), then render only
        while time.perf_counter() - t0 < T:
            if guided:
                training = (time.perf_counter() - t0) < train_until
                it.guiding_mode(training, k > 0)
            it.progression(spp * k, spp)
            if guided and training:
                it.train_fused(4)
            k += 1
        el = time.perf_counter() - t0
        img = it.develop()
        st = it.stats()
        key = "guided" if guided else "unguided"
        out[key] = {"relMSE": relmse(img, ref), "spp": spp * k, "seconds": el, "mpaths_per_s": st["paths"] / el / 1e6,
                    "cells": st.get("guide_cells", 0)}
        it.close()
    out["relMSE_ratio_unguided_over_guided"] = out["unguided"]["relMSE"] / out["guided"]["relMSE"]
    print(json.dumps(out))
