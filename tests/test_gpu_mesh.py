"""Config C4 at test size: the procedural heightfield mesh (vertex normals, roughconductor / roughplastic quadrants, two area
lights) with n = 317 -> 199 712 triangles, CUDA path against the oracle.

  * hit records on camera rays, bounce rays and shadow rays for BOTH traversal kernels (batches of 32 and the persistent
    speculative kernel with per-lane refill): identical primitive, t within 2e-6, (u, v) within 2e-5;
  * a chord set in the pattern of the reference's own kd-tree test/benchmark (src/tests/test_kd.cpp:86-133: uniformly
    random chords through the bounding sphere of a mesh, closest- and any-hit), checked against the oracle's kd-tree;
  * per-sample radiance through the mesh materials, sample by sample;
  * ray / path counters of a whole progression against the oracle's render of the same samples.
The 10 M-triangle scene itself (BASELINE config 3) runs in bench.py's `workloads` block and, through size-independent properties, in
test_full_size_c4_properties below.
"""
import numpy as np
import pytest

from test_gpu_parity import _check_hits, _secondary_rays

pytestmark = pytest.mark.gpu

N_MESH = 317
# triangle edges are 2 / (N_MESH - 1) = 6.3e-3 long: (u, v) carry 1 / 6.3e-3 = 158 times the absolute error of unit-size primitives;
# the hit POINT they interpolate stays within 2e-5 * 6.3e-3 * 158 = 2e-5 of the oracle's
UV_SCALE = (N_MESH - 1) / 2.0


@pytest.fixture(scope="module")
def api(pkg):
    from b200pg import api as _api

    return _api


@pytest.fixture(scope="module")
def mesh(pkg, api, oracle):
    sb = pkg.scenes.mesh_scene(256, 256, n=N_MESH)
    p = api.default_params()
    p.max_depth = 8
    return sb, oracle.scene(sb), api.Integrator(api.Scene.from_builder(sb), p), p


def _sphere_chords(rng, n, center, radius):
    """test_kd.cpp:100-112: two uniform points on the bounding sphere, ray through both."""
    def on_sphere(k):
        v = rng.randn(k, 3)
        return center + radius * v / np.linalg.norm(v, axis=1, keepdims=True)
    a, b = on_sphere(n), on_sphere(n)
    d = b - a
    length = np.linalg.norm(d, axis=1, keepdims=True)
    d /= length
    return np.concatenate([a, np.zeros((n, 1)), d, length], 1).astype(np.float32)


@pytest.mark.parametrize("spec", [0, 3])
def test_mesh_hits_both_traversal_kernels(mesh, spec):
    sb, osc, it, _ = mesh
    it.set_option("trace_spec", spec)
    try:
        rng = np.random.RandomState(11)
        pos = (rng.rand(200000, 2) * [sb.width, sb.height]).astype(np.float32)
        rays = osc.camera_rays(pos)
        tuv_o, prim_o, _ = osc.trace(rays)
        tuv_g, prim_g = it.k_trace(rays)
        assert (prim_o != 0xFFFFFFFF).mean() > 0.3
        # a ray through a shared edge / vertex of two triangles may pick either neighbour: 1 in 10^4
        _check_hits(tuv_o, prim_o, tuv_g, prim_g, max_mismatch=int(1e-4 * len(prim_o)), uv_scale=UV_SCALE)
        r2 = _secondary_rays(rays, tuv_o, prim_o, rng, osc)
        tuv_o2, prim_o2, _ = osc.trace(r2)
        tuv_g2, prim_g2 = it.k_trace(r2)
        # bounce rays graze the neighbouring triangles of the sheet: median error 1e-7, 99.9 % quantile ~2e-5
        _check_hits(tuv_o2, prim_o2, tuv_g2, prim_g2, max_mismatch=int(3e-4 * len(prim_o2)), uv_scale=UV_SCALE, t_scale=20.0)
        assert np.median(np.abs(tuv_o2[:, 0] - tuv_g2[:, 0])[(prim_o2 != 0xFFFFFFFF) & (prim_o2 == prim_g2)]) <= 2e-6
        r3 = r2.copy()
        r3[:, 7] = rng.rand(r3.shape[0]).astype(np.float32) * 2.0
        _, occ_o, _ = osc.trace(r3, shadow=True)
        _, occ_g = it.k_trace(r3, shadow=True)
        assert ((occ_o != 0xFFFFFFFF) != (occ_g != 0xFFFFFFFF)).sum() <= int(3e-4 * len(occ_o))
    finally:
        it.set_option("trace_spec", 3)


@pytest.mark.parametrize("spec", [0, 3])
def test_mesh_chord_set_like_test_kd(mesh, spec):
    sb, osc, it, _ = mesh
    it.set_option("trace_spec", spec)
    try:
        rng = np.random.RandomState(5)
        chords = _sphere_chords(rng, 300000, np.array([0.0, 0.0, 0.0]), 1.5)
        tuv_o, prim_o, _ = osc.trace(chords)
        tuv_g, prim_g = it.k_trace(chords)
        assert 0.05 < (prim_o != 0xFFFFFFFF).mean() < 0.95
        _check_hits(tuv_o, prim_o, tuv_g, prim_g, max_mismatch=int(1e-4 * len(prim_o)), uv_scale=UV_SCALE)
        _, occ_o, _ = osc.trace(chords, shadow=True)
        _, occ_g = it.k_trace(chords, shadow=True)
        assert ((occ_o != 0xFFFFFFFF) != (occ_g != 0xFFFFFFFF)).sum() <= int(1e-4 * len(occ_o))
    finally:
        it.set_option("trace_spec", 3)


def test_mesh_grazing_rays_cooperative_tail(mesh, api):
    """Rays that skim the terrain need many node visits: the traversal kernels hand a ray to the warp-cooperative kernel
    (kernels.cu: k_trace_tail) once it has used up its visit budget. Closest- and any-hit results must be the oracle's with the
    default budget (96 visits: a few per cent of this set), with a budget of 4 (practically every ray is finished by a warp) and
    with the mechanism off."""
    sb, osc, it, _ = mesh
    rng = np.random.RandomState(17)
    n = 60000
    # start on a circle outside the sheet, a little above / below its height range (+-0.05), aim across it almost horizontally
    phi = rng.rand(n) * 2 * np.pi
    o = np.stack([1.6 * np.cos(phi), rng.uniform(-0.04, 0.06, n), 1.6 * np.sin(phi)], 1)
    tgt = np.stack([rng.uniform(-0.9, 0.9, n), rng.uniform(-0.05, 0.05, n), rng.uniform(-0.9, 0.9, n)], 1)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, np.zeros((n, 1)), d, np.full((n, 1), 10.0)], 1).astype(np.float32)
    tuv_o, prim_o, _ = osc.trace(rays)
    _, occ_o, _ = osc.trace(rays, shadow=True)
    assert 0.2 < (prim_o != 0xFFFFFFFF).mean() < 0.98
    try:
        for budget in (-1, 4, 0):  # library default, nearly everything cooperative, off
            it.set_option("tail_visits", budget)
            tuv_g, prim_g = it.k_trace(rays)
            _check_hits(tuv_o, prim_o, tuv_g, prim_g, max_mismatch=int(3e-4 * n), uv_scale=UV_SCALE, t_scale=20.0)
            _, occ_g = it.k_trace(rays, shadow=True)
            assert ((occ_o != 0xFFFFFFFF) != (occ_g != 0xFFFFFFFF)).sum() <= int(3e-4 * n), budget
    finally:
        it.set_option("tail_visits", -1)


def test_mesh_radiance_sample_by_sample(mesh):
    sb, osc, it, p = mesh
    rng = np.random.RandomState(3)
    pix = rng.randint(0, sb.width * sb.height, 60000).astype(np.uint32)
    smp = rng.randint(0, 64, 60000).astype(np.uint32)
    Lo = osc.radiance(p, pix, smp)
    Lg = it.k_radiance(pix, smp)
    err = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-3)
    assert (Lo.max(1) > 0).mean() > 0.2
    # microfacet lobes at grazing angles amplify a 1-ulp difference in the half vector: 0.5 % of the samples
    assert (err > 1e-3).mean() < 5e-3, "fraction of samples off by > 1e-3 relative: %g" % (err > 1e-3).mean()
    assert abs(Lg.mean() - Lo.mean()) <= 2e-3 * Lo.mean()


def test_mesh_radiance_does_not_depend_on_the_schedule(mesh):
    """The hit / miss partition of the shade queue (off / forced on every bounce >= 1) and the visit budget (off / 4 = nearly every
    ray through the warp-cooperative kernel) change the ORDER in which paths are shaded and rays are traversed, never a sample's
    value: radiance per (pixel, sample) must be bit-identical under the partition and equal up to exact-t ties under the budget."""
    sb, osc, it, p = mesh
    rng = np.random.RandomState(23)
    pix = rng.randint(0, sb.width * sb.height, 80000).astype(np.uint32)
    smp = rng.randint(0, 64, 80000).astype(np.uint32)
    try:
        it.set_option("partition", 0)
        base = it.k_radiance(pix, smp)
        it.set_option("partition", 2)
        assert np.array_equal(base, it.k_radiance(pix, smp))
        it.set_option("tail_visits", 4)
        got = it.k_radiance(pix, smp)
        assert (np.abs(got - base).max(1) > 1e-6 * (1 + np.abs(base).max(1))).mean() < 1e-3
        it.set_option("tail_visits", 0)
        got = it.k_radiance(pix, smp)
        assert (np.abs(got - base).max(1) > 1e-6 * (1 + np.abs(base).max(1))).mean() < 1e-3
    finally:
        it.set_option("partition", 1)
        it.set_option("tail_visits", -1)


def test_mesh_progression_counters_and_image(mesh, api):
    sb, osc, it, p = mesh
    it.film_clear()
    s0 = it.stats()
    it.progression(0, 2)
    s1 = it.stats()
    film_o, st = osc.render(p, 0, 2)
    from oracle_lib import develop

    paths = s1["paths"] - s0["paths"]
    assert paths == st["paths"] == sb.width * sb.height * 2
    for key in ("normal_rays", "shadow_rays"):
        g, o = s1[key] - s0[key], st[key]
        assert abs(g - o) <= 2e-3 * o, (key, g, o)
    img_g, img_o = it.develop(), develop(film_o)
    num = np.abs(img_g - img_o).sum()
    assert num / np.abs(img_o).sum() < 5e-3  # same samples on both sides: differences are flipped decisions only


def test_mesh_hits_wide_tree(pkg, api, oracle, monkeypatch):
    """The optional 8-ary quantised tree (B200PG_WIDE_MIN_PRIMS / set_option("wide_bvh"); off by default -- DESIGN.md, measured
    slower than the binary tree): same leaves, conservative quantised boxes, so the hits must be the binary tree's. Includes
    axis-aligned rays (zero direction components: infinite reciprocals are clamped in the wide test)."""
    monkeypatch.setenv("B200PG_WIDE_MIN_PRIMS", "1000")
    sb = pkg.scenes.mesh_scene(256, 256, n=N_MESH)
    p = api.default_params()
    p.max_depth = 8
    it = api.Integrator(api.Scene.from_builder(sb), p)
    monkeypatch.delenv("B200PG_WIDE_MIN_PRIMS")
    osc = oracle.scene(sb)
    rng = np.random.RandomState(21)
    chords = _sphere_chords(rng, 200000, np.array([0.0, 0.0, 0.0]), 1.5)
    down = np.zeros((20000, 8), np.float32)
    down[:, 0], down[:, 2] = rng.rand(20000) * 2 - 1, rng.rand(20000) * 2 - 1
    down[:, 1], down[:, 5], down[:, 7] = 1.0, -1.0, np.inf
    rays = np.concatenate([chords, down])
    tuv_o, prim_o, _ = osc.trace(rays)
    res = {}
    for wide in (0, 1):
        it.set_option("wide_bvh", wide)
        it.set_option("trace_spec", 7)
        tuv_g, prim_g = it.k_trace(rays)
        _check_hits(tuv_o, prim_o, tuv_g, prim_g, max_mismatch=int(2e-4 * len(prim_o)), uv_scale=UV_SCALE)
        _, occ_g = it.k_trace(rays, shadow=True)
        res[wide] = (tuv_g, prim_g, occ_g != 0xFFFFFFFF)
    same = res[0][1] == res[1][1]
    assert (~same).sum() <= 20                                   # rays through a shared edge may pick either neighbour
    assert np.array_equal(res[0][0][same], res[1][0][same])      # ... everything else is bit-identical
    assert (res[0][2] != res[1][2]).sum() <= 5
    # and a whole progression through the wide tree gives the binary tree's image
    imgs = []
    for wide in (0, 1):
        it.set_option("wide_bvh", wide)
        it.film_clear()
        it.progression(0, 2)
        imgs.append(it.develop())
    assert np.abs(imgs[0] - imgs[1]).sum() / np.abs(imgs[0]).sum() < 1e-3


def _terrain_height(x, z, seed=1337, amp=0.05):
    """The analytic surface scenes.heightfield_mesh samples (same seeded sum of six sines), in float64."""
    rng = np.random.RandomState(seed)
    y = np.zeros_like(x, dtype=np.float64)
    for _ in range(6):
        fx, fz = rng.uniform(2, 14, 2)
        ph = rng.uniform(0, 2 * np.pi)
        a = rng.uniform(0.3, 1.0)
        y += a * np.sin(fx * x + fz * z + ph)
    return y * amp / 3.0


def test_full_size_c4_properties(pkg, api):
    """BASELINE config C4 at full size -- 10.0 M triangles, 2048 x 2048 -- through properties that do not need the oracle (its kd-tree
    build alone takes 46 s): closest hits lie on the ANALYTIC terrain the mesh samples (within the linear-interpolation error of a
    0.9 mm grid), any-hit and closest-hit queries agree on which chords are blocked, a sample's radiance does not depend on the
    traversal / shading schedule (hit / miss partition bit for bit, visit budget + warp-cooperative kernel up to exact-t ties), and
    a progression deposits every camera sample in the film."""
    sb = pkg.scenes.mesh_scene(2048, 2048)
    p = api.default_params()
    p.max_depth = 8
    it = api.Integrator(api.Scene.from_builder(sb), p)
    rng = np.random.RandomState(43)
    n = 300000
    chords = _sphere_chords(rng, n, np.array([0.0, 0.0, 0.0]), 1.5)
    # a third of the set skims the sheet: long traversals, the visit budget and the cooperative kernel at work
    k = n // 3
    phi = rng.rand(k) * 2 * np.pi
    o = np.stack([1.6 * np.cos(phi), rng.uniform(-0.04, 0.06, k), 1.6 * np.sin(phi)], 1)
    tgt = np.stack([rng.uniform(-0.9, 0.9, k), rng.uniform(-0.05, 0.05, k), rng.uniform(-0.9, 0.9, k)], 1)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    chords[:k] = np.concatenate([o, np.zeros((k, 1)), d, np.full((k, 1), 10.0)], 1).astype(np.float32)
    tuv, prim = it.k_trace(chords)
    hit = prim != 0xFFFFFFFF
    assert 0.1 < hit.mean() < 0.95
    pt = chords[:, 0:3].astype(np.float64) + tuv[:, 0:1].astype(np.float64) * chords[:, 4:7].astype(np.float64)
    on_sheet = hit & (np.abs(pt[:, 1]) < 0.2)          # the two lights hang at y = 0.9
    assert on_sheet.sum() > 0.9 * hit.sum()
    dy = np.abs(pt[on_sheet, 1] - _terrain_height(pt[on_sheet, 0], pt[on_sheet, 2]))
    # vertical distance to the analytic surface: facet interpolation error <= ~1e-6 plus the ray parameter's rounding, which a
    # grazing ray turns into height error ~ 1e-6 * slope of approach -- bound the bulk tightly and the grazing tail loosely
    assert np.quantile(dy, 0.99) <= 1e-5 and dy.max() <= 2e-4, (np.quantile(dy, 0.99), dy.max())
    assert np.abs(pt[on_sheet][:, [0, 2]]).max() <= 1.0 + 1e-5
    _, occ = it.k_trace(chords, shadow=True)
    assert ((occ != 0xFFFFFFFF) != hit).sum() <= int(2e-4 * n)
    # the same queries with the visit budget off: identical up to exact-t ties on shared edges
    try:
        it.set_option("tail_visits", 0)
        tuv0, prim0 = it.k_trace(chords)
    finally:
        it.set_option("tail_visits", -1)
    assert (prim0 != prim).sum() <= int(2e-4 * n)
    same = (prim0 == prim) & hit
    assert np.abs(tuv0[same, 0] - tuv[same, 0]).max() <= 1e-5
    # radiance per sample under different schedules
    pix = rng.randint(0, sb.width * sb.height, 100000).astype(np.uint32)
    smp = rng.randint(0, 16, 100000).astype(np.uint32)
    try:
        it.set_option("partition", 0)
        base = it.k_radiance(pix, smp)
        it.set_option("partition", 2)
        assert np.array_equal(base, it.k_radiance(pix, smp))
    finally:
        it.set_option("partition", 1)
    assert np.isfinite(base).all() and (base >= 0).all() and (base.max(1) > 0).mean() > 0.1
    # two progressions (the second one takes the adaptive partition): every camera sample ends up in the film
    it.film_clear()
    s0 = it.stats()
    it.progression(0, 1)
    it.progression(1, 1)
    s1 = it.stats()
    f = it.film()
    assert s1["paths"] - s0["paths"] == 2 * 2048 * 2048
    assert np.isfinite(f).all() and (f >= 0).all()
    assert 0.97 <= f[..., 4].sum() / (2 * 2048 * 2048) <= 1.0001
    it.close()
