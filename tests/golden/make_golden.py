"""Generates the committed golden fixtures tests/golden/*.npz.

What they are: REGRESSION vectors of this repo's CPU oracle (oracle/, the restatement of the reference algorithm) on small
seeded inputs -- inputs and outputs side by side -- so that (i) a change of the oracle shows up as a diff against committed
numbers (tests/test_golden.py, CPU) and (ii) the CUDA path can be checked on the GPU box against numbers that were fixed
when the oracle was pinned (tests/test_golden.py, -m gpu).
What they are NOT: outputs of the reference. Those -- for the very same inputs -- are tests/golden/upstream.npz, written by
make_upstream.py from the reference's own code compiled into oracle/_ref (DESIGN.md "Reference build status"), and
tests/test_upstream.py holds these fixtures and the CUDA path to them. The reference's own known answers for this path are
restated in tests/test_oracle_kd_film.py (test_kd.cpp:34-83 clipping vectors, test_dgeom.cpp), tests/test_oracle_bsdf.py
(test_chisquare / test_bsdf.xml) and tests/test_oracle_medium.py.

usage: python tests/golden/make_golden.py        (needs only the oracle: no GPU, no /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from bsdf_cases import bsdf_scene, random_dirs  # noqa: E402
from conftest import load_package  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

SEED = 1337  # the sampler default of the reference (independent.cpp:58)


def params(pkg, **kw):
    p = pkg._abi.default_params()
    p.max_depth = 8
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def cases(pkg):
    """The scenes of the fixtures (shared with tests/test_golden.py)."""
    S = pkg.scenes
    return dict(cornell=lambda: S.cornell_box(64, 64, spp=4), caustic=lambda: S.cornell_caustic(64, 64, spp=4),
                medium=lambda: S.cornell_medium(64, 64, spp=4, res=16))


def secondary_rays(rays, tuv, prim, rng):
    hit = prim != 0xFFFFFFFF
    P = rays[hit, :3] + rays[hit, 4:7] * tuv[hit, 0:1]
    d = random_dirs(rng, P.shape[0])
    back = (d * rays[hit, 4:7]).sum(1) > 0
    d[back] *= -1
    return np.concatenate([P, np.full((P.shape[0], 1), 1e-4, np.float32), d, np.full((P.shape[0], 1), np.inf, np.float32)],
                          1).astype(np.float32)


def generate(pkg, orc):
    """All fixtures as {file stem: {array name: array}} (same seeded stream every time)."""
    C = cases(pkg)
    rng = np.random.RandomState(SEED)
    out = {}

    # ---- traversal + hit records, camera rays, per-sample radiance, film --------------------------------------------
    for name in ("cornell", "caustic"):
        sb = C[name]()
        osc = orc.scene(sb)
        pos = (rng.rand(4096, 2) * [sb.width, sb.height]).astype(np.float32)
        rays = osc.camera_rays(pos)
        tuv, prim, _ = osc.trace(rays)
        r2 = secondary_rays(rays, tuv, prim, rng)[:3000]
        tuv2, prim2, _ = osc.trace(r2)
        r3 = r2.copy()
        r3[:, 7] = rng.rand(r3.shape[0]).astype(np.float32) * 3.0
        _, occ, _ = osc.trace(r3, shadow=True)
        pix = rng.randint(0, sb.width * sb.height, 4096).astype(np.uint32)
        smp = rng.randint(0, 1000, 4096).astype(np.uint32)
        rad = osc.radiance(params(pkg), pix, smp)
        out[name] = dict(pos=pos, rays=rays, tuv=tuv, prim=prim, rays2=r2, tuv2=tuv2, prim2=prim2, rays3=r3,
                         occluded=(occ != 0xFFFFFFFF), pixel=pix, sample=smp, radiance=rad)
        if name == "cornell":
            spos = (rng.rand(20000, 2) * [sb.width, sb.height]).astype(np.float32)
            spos[:32] = np.floor(spos[:32])
            spos[32:64, 0] = 0.0
            srgb = (rng.rand(20000, 3) * 4).astype(np.float32)
            out[name].update(splat_pos=spos, splat_rgb=srgb, film=osc.film_splat(spos, srgb))
            film8, st = osc.render(params(pkg), 0, 2)
            out[name].update(render_film=film8, render_counters=np.array(
                [st["paths"], st["normal_rays"], st["shadow_rays"], st["path_length_sum"]], np.uint64))

    # ---- BSDFs (the parameterisations of data/tests/test_bsdf.xml) -----------------------------------------------------
    sb, idx = bsdf_scene()
    osc = orc.scene(sb)
    bs = {}
    for name, i in idx.items():
        wi, wo, u = random_dirs(rng, 1500), random_dirs(rng, 1500), rng.rand(1500, 2).astype(np.float32)
        o = osc.bsdf(i, wi, wo, u)
        bs[name] = dict(wi=wi, wo_in=wo, u=u, **o)
    out["bsdf"] = {"%s/%s" % (n, k): v for n, d in bs.items() for k, v in d.items()}

    # ---- medium: grid look-up, free-flight sampling / transmittance ----------------------------------------------------
    sb = C["medium"]()
    osc = orc.scene(sb)
    p = (rng.rand(4096, 3) * 2.4 - [1.2, 0.2, 1.2]).astype(np.float32)
    dens = osc.grid_lookup(0, p)
    o = (rng.rand(2048, 3) * [1.6, 1.6, 1.6] + [-0.8, 0.2, -0.8]).astype(np.float32)
    d = random_dirs(rng, 2048)
    mr = np.concatenate([o, np.zeros((2048, 1), np.float32), d, (rng.rand(2048, 1) * 2 + 0.1).astype(np.float32)], 1).astype(np.float32)
    t, tr, wo, pdf = osc.medium_sample(0, mr)
    out["medium"] = dict(points=p, density=dens, rays=mr, t=t, transmittance=tr, wo=wo, pdf=pdf)

    # ---- guiding field: query, binning, E-step statistics, a whole training update ---------------------------------------
    sb = C["caustic"]()
    osc = orc.scene(sb)
    K = 8
    fld = orc.field(K, (0, 0, 0), (1, 1, 1))
    gp = params(pkg, guiding=1, guide_max_components=K, guide_max_cell_samples=3000)
    sink = orc.samples()
    for k in range(3):  # three oracle training updates -> a small tree with fitted mixtures
        sink.clear()
        osc.render(gp, 4 * k, 4, field=fld if k else None, sink=sink)
        fld.train_sink(sink, 4, 3000.0)
    snap = fld.snapshot()
    n = 4096
    qpos = (rng.rand(n, 3) * [2.2, 2.2, 2.2] - [1.1, 0.1, 1.1]).astype(np.float32)
    qdir, qu = random_dirs(rng, n), rng.rand(n, 3).astype(np.float32)
    q = fld.pdf_sample(qpos, qdir, qu)
    cell, perm, off = fld.bin(qpos)
    sink.clear()
    osc.render(gp, 100, 2, field=fld, sink=sink)
    s = sink.get()
    m = min(20000, s["pos"].shape[0])
    s = {k: np.ascontiguousarray(v[:m]) for k, v in s.items()}
    stats = fld.estep(s)
    fld.train(s, 4, 3000.0)
    out["guiding"] = dict(K=np.array([K], np.uint32), field=snap, qpos=qpos, qdir=qdir, qu=qu, q_pdf=q["pdf"], q_dir=q["dir"],
                          q_spdf=q["spdf"], q_cell=q["cell"], bin_cell=cell, bin_perm=perm, bin_offsets=off,
                          s_pos=s["pos"], s_dir=s["dir"], s_weight=s["weight"], s_pdf=s["pdf"], s_dist=s["dist"],
                          estep_stats=stats, field_after=fld.snapshot())

    return out


def main():
    out = generate(load_package(), Oracle())
    for name, d in out.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print("%-10s %7.1f KB  %s" % (name, os.path.getsize(path) / 1024, ", ".join(sorted(d)[:6]) + (" ..." if len(d) > 6 else "")))


if __name__ == "__main__":
    main()
