"""Live pin of the oracle port against the reference's own code (oracle/_ref, DESIGN.md "Reference build status"): where
tests/test_upstream.py compares committed outputs for the golden inputs, this file runs both sides on more scenes and on the
integrator's parameter space, sample by sample (the replay sampler hands ProgressiveMIPathTracer::Li the counter-based stream
the port uses). Skipped where oracle/_ref is not built."""
import numpy as np
import pytest

import ref_lib
from transport_cases import form_factor_scene, furnace_scene

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built (no /root/reference here)")


def both(pkg, oracle, sb):
    osc = oracle.scene(sb)
    return osc, ref_lib.RefScene(desc=osc.desc, keep=osc._keep)


def params(pkg, **kw):
    p = pkg._abi.default_params()
    p.max_depth = 8
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def compare_li(pkg, osc, rs, p, width, height, spp=4, frac=1e-3):
    pix = np.repeat(np.arange(width * height, dtype=np.uint32), spp)
    smp = np.tile(np.arange(spp, dtype=np.uint32), width * height)
    lo, (lr, _) = osc.radiance(p, pix, smp), rs.radiance(p, pix, smp)
    rel = np.abs(lo - lr).max(1) / np.maximum(np.abs(lr).max(1), 1e-3)
    assert (rel > 1e-4).mean() <= frac, (rel > 1e-4).mean()   # flipped discrete decisions (rounding on two compilers)
    assert (rel > 1e-2).mean() <= frac / 4
    assert abs(lo.mean() - lr.mean()) <= 1e-3 * abs(lr.mean()) + 1e-7
    return lr


SCENES = {
    "cornell": lambda S: S.cornell_box(48, 48, spp=4),
    "caustic": lambda S: S.cornell_caustic(48, 48, spp=4),
    "mesh": lambda S: S.mesh_scene(48, 48, spp=4, n=40),
}


@pytest.mark.parametrize("scene", sorted(SCENES))
@pytest.mark.parametrize("kw", [dict(), dict(max_depth=1), dict(max_depth=2), dict(max_depth=3), dict(max_depth=-1, rr_depth=2),
                                dict(strict_normals=1), dict(hide_emitters=1), dict(use_nee=0)],
                         ids=lambda kw: ",".join("%s=%s" % kv for kv in kw.items()) or "defaults")
def test_li_matches_the_reference_sample_by_sample(pkg, oracle, scene, kw):
    sb = SCENES[scene](pkg.scenes)
    osc, rs = both(pkg, oracle, sb)
    compare_li(pkg, osc, rs, params(pkg, **kw), sb.width, sb.height)


@pytest.mark.parametrize("light", ["rectangle", "trimesh"])
def test_area_lights_and_default_bsdfs_match_the_reference(pkg, oracle, light):
    """The form-factor scene: a shape without a BSDF that carries an emitter (Shape::configure gives it a black diffuse BSDF,
    shape.cpp:48-70), as a rectangle (Rectangle::samplePosition) and as a two-triangle mesh (TriMesh::samplePosition over the
    area distribution, trimesh.cpp:412-423); with and without next-event estimation, against the closed form as well."""
    sb, centre, want = form_factor_scene(pkg, light)
    osc, rs = both(pkg, oracle, sb)
    for nee in (1, 0):
        compare_li(pkg, osc, rs, params(pkg, max_depth=2, use_nee=nee), sb.width, sb.height, spp=64)
    n = 40000
    pix, smp = np.full(n, centre, np.uint32), np.arange(n, dtype=np.uint32)
    lr = rs.radiance(params(pkg, max_depth=2), pix, smp)[0].astype(np.float64)
    assert np.all(np.abs(lr.mean(0) - want) <= 4 * lr.std(0) / np.sqrt(n) + 2e-3 * want)  # the reference itself hits rho L F


@pytest.mark.parametrize("glass", [False, True])
def test_furnace_matches_the_reference(pkg, oracle, glass):
    """Six emitting, reflecting walls (twosided is not involved: rectangles seen from inside), optionally a glass cube: unbounded
    depth with Russian roulette; the reference's own mean is L / (1 - rho)."""
    sb, want = furnace_scene(pkg, glass=glass)
    osc, rs = both(pkg, oracle, sb)
    lr = compare_li(pkg, osc, rs, params(pkg, max_depth=-1, rr_depth=5), sb.width, sb.height, spp=64, frac=3e-3).astype(np.float64)
    assert abs(lr.mean() - want) <= 4 * lr.std() / np.sqrt(lr.size) + 2e-3 * want


def test_every_bsdf_of_the_path_in_one_scene(pkg, oracle):
    from bsdf_cases import bsdf_scene

    sb, idx = bsdf_scene()
    osc, rs = both(pkg, oracle, sb)
    compare_li(pkg, osc, rs, params(pkg), sb.width, sb.height, spp=64, frac=2e-3)


@pytest.mark.parametrize("axis", [0, 1, 2, 3, 4])
def test_camera_rays_for_every_fov_axis(pkg, oracle, axis):
    """perspective.cpp:126-180: fovAxis x / y / diagonal / smaller / larger on a non-square film."""
    sb = pkg.scenes.cornell_box(96, 40, spp=1)
    sb.sensor["fov_axis"] = axis
    osc, rs = both(pkg, oracle, sb)
    pos = (np.random.RandomState(axis).rand(2000, 2) * [96, 40]).astype(np.float32)
    np.testing.assert_allclose(osc.camera_rays(pos), rs.camera_rays(pos), rtol=1e-6, atol=2e-6)


def test_kd_tree_hits_on_random_chords_of_a_mesh(pkg, oracle):
    """The pattern of src/tests/test_kd.cpp:86-133 (random chords of the bounding sphere through a mesh): closest hits and
    shadow rays of the reference's SAH kd-tree (gkdtree.h build, sahkdtree3.h traversal, TriAccel) against the port's."""
    sb = pkg.scenes.mesh_scene(32, 32, spp=1, n=120)   # 28 k triangles
    osc, rs = both(pkg, oracle, sb)
    rng = np.random.RandomState(11)
    a, b = rng.randn(20000, 3), rng.randn(20000, 3)
    a, b = 1.6 * a / np.linalg.norm(a, axis=1, keepdims=True), 1.6 * b / np.linalg.norm(b, axis=1, keepdims=True)
    c = np.array([0.0, 0.3, 0.0])
    d = b - a
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([a + c, np.zeros((20000, 1)), d, np.full((20000, 1), np.inf)], 1).astype(np.float32)
    tuv, prim, _ = osc.trace(rays)
    hit = rs.intersect(rays)
    assert (prim != hit["prim"]).sum() <= 2 and (prim != 0xFFFFFFFF).mean() > 0.2
    m = (prim == hit["prim"]) & (prim != 0xFFFFFFFF)
    assert np.abs(tuv[m, 0] - hit["t"][m]).max() <= 2e-5
    _, occ, _ = osc.trace(rays, shadow=True)
    assert ((occ != 0xFFFFFFFF) != rs.occluded(rays)).sum() <= 2


def test_two_lights_with_different_sampling_weights(pkg, oracle):
    """Scene::sampleEmitterDirect / pdfEmitterDirect over the discrete emitter distribution (scene.cpp:871-895, pmf.h:124-188: the
    sample is re-used after the choice) with unequal samplingWeight."""
    S = pkg.scenes
    sb = S.SceneBuilder(24, 24, spp=1)
    X = (1, 0, 0)
    sb.rectangle([S.scale(50, 50, 1), S.rotate(X, -90.0)], bsdf=sb.diffuse((0.5, 0.5, 0.5)))
    for (hx, hz), c, L, w in [((0.4, 0.3), (0.0, 1.2, 0.0), (4.0, 4.0, 4.0), 1.0), ((0.2, 0.6), (0.9, 2.0, -0.5), (1.0, 9.0, 2.0), 3.5)]:
        sb.rectangle([S.scale(hx, hz, 1), S.rotate(X, 90.0), S.translate(*c)], bsdf=-1, radiance=L)
        sb.emitters[-1]["weight"] = w
    sb.set_camera((3.0, 1.0, 2.5), (0.2, 0.0, 0.1), (0, 1, 0), 30.0)
    osc, rs = both(pkg, oracle, sb)
    compare_li(pkg, osc, rs, params(pkg, max_depth=3), sb.width, sb.height, spp=32)
    u = np.random.RandomState(6).rand(4000, 2).astype(np.float32)
    refp, refn = np.array([0.2, 0.0, 0.1], np.float32), np.array([0, 1, 0], np.float32)
    a, b = osc.emitter_sample(refp, refn, u), rs.emitter_sample(refp, refn, u)
    np.testing.assert_allclose(a[0], b[0], atol=3e-6)            # directions: the same light was chosen for every sample
    np.testing.assert_allclose(a[2], b[2], rtol=2e-5)            # pdf includes the choice probability
    np.testing.assert_allclose(a[3], b[3], rtol=2e-5, atol=1e-6)


def test_kd_tree_build_and_traversal_counters(pkg, oracle):
    """SURVEY.md 8(d) takes the algorithmic bytes per ray from the node visits / index reads of a counting kd traversal and
    names the reference's own (rayIntersectHavranCollectStatistics, sahkdtree3.h:330-429). Here that method runs on the
    reference's tree next to the port's counters on the port's tree, same rays, and the two SAH builds are compared node by
    node (depth-first dump: axis, split plane, leaf sizes): the port's tree IS the reference's tree -- C1, C2, the media scene,
    a 28 k-triangle mesh (exact O(n log n) stage with perfect splits) and a 200 k-triangle mesh (min-max binning stage, the
    hand-over to the exact stage, the parallel build's no-retraction rule) -- and the counters are equal to the last visit.
    (Getting there took the reference's order of operations in the SAH probabilities without fused multiply-adds -- equal-cost
    candidate planes abound on a regular mesh and the strict `<` keeps the first -- its tie rule for planar primitives, its
    retraction of splits that did not pay off, its depth count from 1, and min-max binning over the tight bounds with the
    partition by bin index.)"""
    import ctypes as C
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg

    fp = C.POINTER(C.c_float)
    oracle.lib.orc_kd_dump.argtypes = [C.c_void_p, fp, C.c_int]
    ref_lib.lib().ref_kd_dump.argtypes = [C.c_void_p, fp, C.c_int]
    S = pkg.scenes
    for name, sb in (("cornell", S.cornell_box(64, 64, 4)), ("caustic", S.cornell_caustic(64, 64, 4)),
                     ("medium", S.cornell_medium(32, 32, 4, res=8)), ("mesh 28 k", S.mesh_scene(64, 64, 4, n=120)),
                     ("mesh 200 k", S.mesh_scene(64, 64, 4, n=317))):
        osc, rs = both(pkg, oracle, sb)
        N = 400000
        a, b = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
        na = oracle.lib.orc_kd_dump(osc.h, a.ctypes.data_as(fp), N)
        nb = ref_lib.lib().ref_kd_dump(rs.h, b.ctypes.data_as(fp), N)
        assert na == nb == osc.kd_info()["nodes"], (name, na, nb)
        assert np.array_equal(a[:na], b[:nb]), name
        rng = np.random.RandomState(1)
        rays = osc.camera_rays((rng.rand(6000, 2) * [sb.width, sb.height]).astype(np.float32))
        tuv, prim, _ = osc.trace(rays)
        rays = np.concatenate([rays, mg.secondary_rays(rays, tuv, prim, rng)])
        _, _, c = osc.trace(rays)
        k = rs.kd_count(rays)
        assert (c["nodes"] - c["leaves"], c["indices"]) == (k["inner"], k["indices"]), (name, c, k)
