"""The closed-form scenes of tests/transport_cases.py through the CUDA path (b200pg_k_radiance): form-factor direct
illumination, the furnace (plain, glass cube, scattering medium). Prints mean +- standard error next to the expected value.
usage (GPU box): python tools/gpu_analytic_check.py            -- about 10 s of GPU time
The volumetric furnace with next-event estimation is EXPECTED to read ~10 % high: the reference quirk documented in
DESIGN.md section 2 / oracle/oracle_volpath.h, reproduced for parity (the 'no NEE' leg is exact)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from b200pg import api  # noqa: E402
from transport_cases import form_factor_scene, furnace_scene  # noqa: E402


def params(**kw):
    p = api.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def run(name, sb, want, pix, n, **kw):
    it = api.Integrator(api.Scene.from_builder(sb), params(**kw))
    rad = it.k_radiance(pix, np.arange(n, dtype=np.uint32)).astype(np.float64)
    mean, sem = rad.mean(0), rad.std(0) / np.sqrt(n)
    dev = np.abs(mean - want) / np.maximum(sem, 1e-12)
    print("%-46s mean %s  +- %s  expected %s  (%.1f sigma)" % (name, np.round(mean, 4), np.round(sem, 4), np.round(want, 4), dev.max()))
    it.close()


def main():
    n = 400000
    rng = np.random.RandomState(0)
    for light in ("rectangle", "trimesh"):
        for nee in (1, 0):
            sb, centre, want = form_factor_scene(pkg, light)
            run("form factor, %s light, NEE %d" % (light, nee), sb, want, np.full(n, centre, np.uint32), n, max_depth=2, use_nee=nee)
    pix = rng.randint(0, 256, n).astype(np.uint32)
    sb, want = furnace_scene(pkg)
    run("furnace", sb, want, pix, n, max_depth=-1, rr_depth=5)
    run("furnace, maxDepth 4 (expect 1.875)", sb, 1.875, pix, n, max_depth=4, rr_depth=100)
    sb, want = furnace_scene(pkg, glass=True)
    run("furnace + glass cube", sb, want, pix, n, max_depth=-1, rr_depth=5)
    run("furnace + glass cube, volumetric integrator", sb, want, pix, n, max_depth=-1, rr_depth=5, volumetric=1)
    for med in (("isotropic", 0.0, "woodcock"), ("hg", 0.7, "woodcock"), ("hg", -0.3, "simpson")):
        sb, want = furnace_scene(pkg, medium=med)
        run("furnace + medium %s, no NEE" % (med,), sb, want, pix, n, max_depth=-1, rr_depth=5, volumetric=1, use_nee=0)
        run("furnace + medium %s, NEE (reference quirk)" % (med,), sb, want, pix, n, max_depth=-1, rr_depth=5, volumetric=1)


if __name__ == "__main__":
    main()
