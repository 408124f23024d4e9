"""Generates tests/golden/upstream.npz: outputs of the REFERENCE ITSELF for the inputs of the golden fixtures.

The reference's own classes (ShapeKDTree, the BSDF / emitter / sensor / film / medium plugins, ProgressiveMIPathTracer::Li,
ProgressiveMonteCarloIntegrator::render) are compiled from /root/reference by oracle/Makefile.ref into oracle/_ref and driven
through oracle/ref_harness/ref_harness.cpp (tests/ref_lib.py). This script feeds them the INPUT arrays of tests/golden/*.npz
(rays, camera-sample ids, BSDF directions, splat positions, medium points -- written by make_golden.py) plus a mesh case of
its own, and stores what the reference returns. The result travels to the GPU box, where neither /root/reference nor a
compiler for it exists:
  * tests/test_upstream.py (CPU) checks the oracle port's fixtures against these numbers -- the pin of the oracle;
  * tests/test_upstream.py (-m gpu) checks the CUDA path against them directly.

usage: python tests/golden/make_upstream.py      (needs oracle/_ref: `make -C oracle -f Makefile.ref` in the build container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402
from bsdf_cases import bsdf_scene, random_dirs  # noqa: E402
from conftest import load_package  # noqa: E402

BSDF_KEYS = ("eval", "pdf", "wo", "weight", "spdf", "flags")
HIT_KEYS = ("t", "p", "geo_n", "sh_n", "prim")


def mesh_case(pkg):
    """A 7 k-triangle height field with vertex normals under rough BSDFs (the C4 scene at test size)."""
    return pkg.scenes.mesh_scene(64, 64, spp=4, n=60)


def mesh_inputs(pkg, rs):
    rng = np.random.RandomState(4242)
    sb = mesh_case(pkg)
    pos = (rng.rand(4096, 2) * [sb.width, sb.height]).astype(np.float32)
    rays = rs.camera_rays(pos)
    hit = rs.intersect(rays)
    tuv = np.stack([hit["t"], np.zeros_like(hit["t"]), np.zeros_like(hit["t"])], 1)
    r2 = mg.secondary_rays(rays, tuv, hit["prim"], rng)[:3000]
    r3 = r2.copy()
    r3[:, 7] = rng.rand(r3.shape[0]).astype(np.float32) * 3.0
    pix = rng.randint(0, sb.width * sb.height, 4096).astype(np.uint32)
    smp = rng.randint(0, 1000, 4096).astype(np.uint32)
    return dict(pos=pos, rays=rays, rays2=r2, rays3=r3, pixel=pix, sample=smp)


def light_cases(pkg):
    """Scenes that exercise the emitter code paths: a rectangle light and the same light as a two-triangle mesh over a diffuse
    floor (shapes without a BSDF of their own), two lights with unequal sampling weights, the furnace with a glass cube."""
    from transport_cases import form_factor_scene, furnace_scene

    S = pkg.scenes

    def two_lights():
        sb = S.SceneBuilder(24, 24, spp=1)
        X = (1, 0, 0)
        sb.rectangle([S.scale(50, 50, 1), S.rotate(X, -90.0)], bsdf=sb.diffuse((0.5, 0.5, 0.5)))
        for (hx, hz), c, L, w in [((0.4, 0.3), (0.0, 1.2, 0.0), (4.0, 4.0, 4.0), 1.0), ((0.2, 0.6), (0.9, 2.0, -0.5), (1.0, 9.0, 2.0), 3.5)]:
            sb.rectangle([S.scale(hx, hz, 1), S.rotate(X, 90.0), S.translate(*c)], bsdf=-1, radiance=L)
            sb.emitters[-1]["weight"] = w
        sb.set_camera((3.0, 1.0, 2.5), (0.2, 0.0, 0.1), (0, 1, 0), 30.0)
        return sb

    return dict(light_rect=lambda: form_factor_scene(pkg, "rectangle")[0], light_mesh=lambda: form_factor_scene(pkg, "trimesh")[0],
                two_lights=two_lights, furnace_glass=lambda: furnace_scene(pkg, glass=True)[0])


def light_samples(sb, n=3000):
    rng = np.random.RandomState(77)
    return rng.randint(0, sb.width * sb.height, n).astype(np.uint32), rng.randint(0, 1000, n).astype(np.uint32)


def generate(pkg, ref_lib, gold):
    """{key: array}: the reference's outputs, keys '<case>/<name>'."""
    out = {}
    C = mg.cases(pkg)
    scenes = dict(cornell=C["cornell"], caustic=C["caustic"], mesh=lambda: mesh_case(pkg))
    for name, make in scenes.items():
        rs = ref_lib.RefScene(make())
        g = gold[name] if name in gold else mesh_inputs(pkg, rs)
        if name == "mesh":
            out.update({"mesh/in_" + k: v for k, v in g.items()})
        np.testing.assert_allclose(rs.camera_rays(g["pos"]), g["rays"], rtol=1e-6, atol=2e-6)  # the inputs are the reference's camera rays
        out[name + "/camera_rays"] = rs.camera_rays(g["pos"])
        for tag, rays in (("", g["rays"]), ("2", g["rays2"])):
            hit = rs.intersect(rays)
            for k in HIT_KEYS:
                out["%s/%s%s" % (name, k, tag)] = hit[k]
        out[name + "/occluded"] = rs.occluded(g["rays3"])
        rad, pos = rs.radiance(mg.params(pkg), g["pixel"], g["sample"])
        out[name + "/radiance"] = rad
        if name == "cornell":
            out[name + "/film"] = rs.film_splat(g["splat_pos"], g["splat_rgb"])
            out[name + "/render_film"] = rs.render(mg.params(pkg), 0, 2)[0]
            out[name + "/render_film_clamped"] = rs.render(mg.params(pkg, max_component_value=0.75), 0, 2)[0]
            # Scene::sampleEmitterDirect / pdfEmitterDirect from a point in the open part of the box
            u = g["splat_rgb"][:2000, :2] / 4.0
            refp, refn = np.array([-0.5, 1.4, 0.6], np.float32), np.array([0, 1, 0], np.float32)
            d, dist, pdf, val = rs.emitter_sample(refp, refn, u)
            out.update({name + "/em_u": u, name + "/em_ref": refp, name + "/em_refn": refn, name + "/em_d": d, name + "/em_dist": dist,
                        name + "/em_pdf": pdf, name + "/em_value": val, name + "/em_pdf_query": rs.emitter_pdf(refp, refn, d)})
    for name, make in light_cases(pkg).items():
        sb = make()
        pix, smp = light_samples(sb)
        out[name + "/radiance"] = ref_lib.RefScene(sb).radiance(mg.params(pkg, max_depth=-1 if name == "furnace_glass" else 3), pix, smp)[0]
    sb, idx = bsdf_scene()
    rs = ref_lib.RefScene(sb)
    for bname, i in idx.items():
        g = {k: gold["bsdf"]["%s/%s" % (bname, k)] for k in ("wi", "wo_in", "u")}
        r = rs.bsdf(i, g["wi"], g["wo_in"], g["u"])
        out.update({"bsdf/%s/%s" % (bname, k): r[k] for k in BSDF_KEYS})
    rs = ref_lib.RefScene(C["medium"]())
    g = gold["medium"]
    out["medium/density"] = rs.grid_lookup(0, g["points"])
    t, ps, tr = rs.medium_sample(0, g["rays"])
    out["medium/t"], out["medium/transmittance"] = t, tr
    rng = np.random.RandomState(99)
    wi, wo, u = random_dirs(rng, 1000), random_dirs(rng, 1000), rng.rand(1000, 2).astype(np.float32)
    ev, swo, pdf = rs.phase(0, wi, wo, u)
    out.update({"medium/ph_wi": wi, "medium/ph_wo_in": wo, "medium/ph_u": u, "medium/ph_eval": ev, "medium/ph_wo": swo, "medium/ph_pdf": pdf})
    return out


def load_gold():
    return {f: dict(np.load(os.path.join(HERE, f + ".npz"))) for f in ("cornell", "caustic", "bsdf", "medium")}


def main():
    pkg = load_package()
    import ref_lib

    if not ref_lib.available():
        raise SystemExit("oracle/_ref/libref_harness.so is missing: make -C oracle -f Makefile.ref (needs /root/reference)")
    out = generate(pkg, ref_lib, load_gold())
    path = os.path.join(HERE, "upstream.npz")
    np.savez_compressed(path, **out)
    print("upstream.npz  %.1f KB  %d arrays" % (os.path.getsize(path) / 1024, len(out)))


if __name__ == "__main__":
    main()
