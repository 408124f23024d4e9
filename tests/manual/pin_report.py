"""Writes profiles/r02_reference_pin.txt: how closely the oracle port follows the REFERENCE ITSELF (oracle/_ref: the reference's
libraries and plugins compiled from /root/reference, driven through oracle/ref_harness with the replay sampler), stage by stage,
as measured numbers rather than pass / fail. The tests that assert these are tests/test_upstream.py, test_ref_pin.py,
test_ref_xml_semantics.py and test_ref_meshes.py.       usage: python tests/manual/pin_report.py   (CPU only, ~1 min)"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "golden"))
import conftest  # noqa: E402

pkg = conftest.load_package()
import make_golden as mg  # noqa: E402
import oracle_lib  # noqa: E402
import ref_lib  # noqa: E402
from bsdf_cases import bsdf_scene, random_dirs  # noqa: E402
from transport_cases import furnace_scene  # noqa: E402

orc = oracle_lib.Oracle()
S, A = pkg.scenes, pkg._abi
out = []


def say(fmt, *a):
    line = fmt % a
    print(line, flush=True)
    out.append(line)


def both(sb):
    o = orc.scene(sb)
    return o, ref_lib.RefScene(desc=o.desc, keep=o._keep)


def params(**kw):
    p = A.default_params()
    p.max_depth = 8
    for k, v in kw.items():
        setattr(p, k, v)
    return p


say("# oracle port vs the compiled reference (oracle/_ref), %s, %d host threads", time.strftime("%Y-%m-%d"), orc.num_threads())
say("# every line: what was compared | n | agreement")
say("")
say("## traversal and hit records (Scene::rayIntersect: gkdtree.h build, sahkdtree3.h traversal, TriAccel / Rectangle, fillIntersectionRecord)")
for name, sb in (("C1 cornell", S.cornell_box(64, 64, 4)), ("C2 caustic", S.cornell_caustic(64, 64, 4)), ("mesh 28 k tris", S.mesh_scene(64, 64, 4, n=120)),
                 ("mesh 200 k tris", S.mesh_scene(64, 64, 4, n=317))):
    o, r = both(sb)
    rng = np.random.RandomState(1)
    rays = o.camera_rays((rng.rand(20000, 2) * 64).astype(np.float32))
    tuv, prim, _ = o.trace(rays)
    rays = np.concatenate([rays, mg.secondary_rays(rays, tuv, prim, rng)])
    tuv, prim, c = o.trace(rays)
    io, ir = o.intersect(rays), r.intersect(rays)
    hit = ir["prim"] != 0xFFFFFFFF
    same = prim == ir["prim"]
    k = r.kd_count(rays)
    say("%-16s | %6d rays | primitive ids differ on %d; |dt| max %.1e; normals max %.1e; shading tangent max %.1e; shadow rays differ on %d; "
        "counting traversal: inner nodes %d vs %d, index entries %d vs %d", name, len(rays), (~same).sum(),
        np.abs(io["t"][hit & same] - ir["t"][hit & same]).max(), np.abs(io["sh_n"][hit & same] - ir["sh_n"][hit & same]).max(),
        np.abs(io["sh_s"][hit & same] - ir["sh_s"][hit & same]).max(),
        ((o.trace(rays, shadow=True)[1] != 0xFFFFFFFF) != r.occluded(rays)).sum(), c["nodes"] - c["leaves"], k["inner"], c["indices"], k["indices"])
say("")
say("## BSDF::eval / pdf / sample (local frame, 20 000 direction pairs each)")
sb, idx = bsdf_scene()
o, r = both(sb)
rng = np.random.RandomState(3)
for name, i in idx.items():
    wi, wo, u = random_dirs(rng, 20000), random_dirs(rng, 20000), rng.rand(20000, 2).astype(np.float32)
    a, b = o.bsdf(i, wi, wo, u), r.bsdf(i, wi, wo, u)
    rel = lambda x, y: np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-3))
    good = b["spdf"] > 0
    w = (np.abs(a["weight"][good] - b["weight"][good]) / np.maximum(np.abs(b["weight"][good]), 1e-3)).max(1)
    say("%-32s | eval rel %.1e, pdf rel %.1e | failed-sample sets equal: %s, lobe flags equal: %s | sampled wo max %.1e | weight rel p99 %.1e, max %.1e",
        name, rel(a["eval"], b["eval"]), rel(a["pdf"], b["pdf"]), np.array_equal(a["spdf"] > 0, good),
        np.array_equal(a["flags"][good], b["flags"][good]), np.abs(a["wo"][good] - b["wo"][good]).max(), np.percentile(w, 99), w.max())
say("")
say("## ProgressiveMIPathTracer::Li per camera sample (replay sampler; 48 x 48 x 16 samples): fraction of samples differing by > 1e-4 / > 1e-2 relative")
for name, make in (("C1 cornell", lambda: S.cornell_box(48, 48, 4)), ("C2 caustic", lambda: S.cornell_caustic(48, 48, 4)), ("mesh", lambda: S.mesh_scene(48, 48, 4, n=40))):
    for kw in (dict(), dict(max_depth=2), dict(max_depth=-1, rr_depth=2), dict(strict_normals=1), dict(hide_emitters=1), dict(use_nee=0)):
        o, r = both(make())
        pix = np.repeat(np.arange(48 * 48, dtype=np.uint32), 16)
        smp = np.tile(np.arange(16, dtype=np.uint32), 48 * 48)
        lo, (lr, _) = o.radiance(params(**kw), pix, smp), r.radiance(params(**kw), pix, smp)
        rel = np.abs(lo - lr).max(1) / np.maximum(np.abs(lr).max(1), 1e-3)
        say("%-10s %-28s | %d samples | %.1e / %.1e | means %.6f vs %.6f", name, ",".join("%s=%s" % kv for kv in kw.items()) or "defaults", len(pix),
            (rel > 1e-4).mean(), (rel > 1e-2).mean(), lo.mean(), lr.mean())
say("")
say("## film: ImageBlock::put and the whole render loop (Scene::preprocess + render on LocalWorkers, Film::put)")
o, r = both(S.cornell_box(64, 64, 4))
rng = np.random.RandomState(5)
pos, rgb = (rng.rand(20000, 2) * 64).astype(np.float32), rng.rand(20000, 3).astype(np.float32)
say("ImageBlock::put, 20 000 splats       | max abs difference %.1e (bit for bit: %s)", np.abs(o.film_splat(pos, rgb) - r.film_splat(pos, rgb)).max(),
    np.array_equal(o.film_splat(pos, rgb), r.film_splat(pos, rgb)))
fo, fr = o.render(params(), 0, 4)[0], r.render(params(), 0, 4)[0]
say("render loop, 64 x 64 x 4 spp         | film max abs difference %.1e on values up to %.0f; relative L1 %.1e", np.abs(fo - fr).max(), fr.max(),
    np.abs(fo - fr).sum() / np.abs(fr).sum())
fo, fr = o.render(params(max_component_value=0.75), 0, 4)[0], r.render(params(max_component_value=0.75), 0, 4)[0]
say("same, maxComponentValue = 0.75       | film max abs difference %.1e", np.abs(fo - fr).max())
say("")
say("## heterogeneous medium (gridvolume 32^3, scale 20)")
for method in ("woodcock", "simpson"):
    sb = S.cornell_medium(48, 48, 4, res=32)
    for m in sb.media:
        m["method"] = A.MEDIUM_SIMPSON if method == "simpson" else A.MEDIUM_WOODCOCK
    o, r = both(sb)
    md = o.desc.media[0]
    lo_, hi_ = np.array(md.aabb_min[:]), np.array(md.aabb_max[:])
    rng = np.random.RandomState(7)
    pts = (lo_ + (hi_ - lo_) * rng.rand(5000, 3)).astype(np.float32)
    a, b = lo_ + (hi_ - lo_) * (0.02 + 0.96 * rng.rand(3000, 3)), lo_ + (hi_ - lo_) * (0.02 + 0.96 * rng.rand(3000, 3))
    d = b - a
    L = np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([a, np.zeros((3000, 1)), d / L, L * 0.98], 1).astype(np.float32)
    to, tro, _, _ = o.medium_sample(0, rays)
    tr, _, trr = r.medium_sample(0, rays)
    fin = np.isfinite(tr)
    say("%-8s | lookupFloat max %.1e | sampleDistance: same interaction / escape decisions: %s, |dt| max %.1e | evalTransmittance max %.1e",
        method, np.abs(o.grid_lookup(0, pts) - r.grid_lookup(0, pts)).max(), np.array_equal(np.isfinite(to), fin), np.abs(to[fin] - tr[fin]).max(),
        np.abs(tro - trr).max())
say("")
say("## ProgressiveVolumetricPathTracer::Li in the mean (furnace with a scattering medium inside, expected 2 without next-event estimation; 100 000 samples)")
for phase, g, method in (("hg", 0.5, "woodcock"), ("isotropic", 0.0, "simpson")):
    sb, want = furnace_scene(pkg, medium=(phase, g, method))
    o, r = both(sb)
    rng = np.random.RandomState(2)
    n = 100000
    pix, smp = rng.randint(0, 256, n).astype(np.uint32), np.arange(n, dtype=np.uint32)
    for nee in (0, 1):
        P = params(max_depth=-1, rr_depth=5, volumetric=1, use_nee=nee)
        ro, rr = o.radiance(P, pix, smp).astype(np.float64).mean(1), r.radiance(P, pix, smp)[0].astype(np.float64).mean(1)
        say("%-9s %-8s useNee=%d | port %.4f +- %.4f | reference %.4f +- %.4f", phase, method, nee, ro.mean(), ro.std() / np.sqrt(n), rr.mean(), rr.std() / np.sqrt(n))

path = os.path.join(conftest.ROOT, "profiles", "r02_reference_pin.txt")
with open(path, "w") as f:
    f.write("\n".join(out) + "\n")
print("wrote", path)
