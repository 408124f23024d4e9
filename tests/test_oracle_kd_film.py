"""Oracle self-checks that do not need a GPU: SAH kd-tree + Havran traversal against brute force (the pattern of the
reference's kd-tree tests, src/tests/test_kd.cpp:86-217), film splatting against a direct numpy restatement of
ImageBlock::put (include/mitsuba/render/imageblock.h:151-197) and the filter table (rfilter.cpp:38-56)."""
import numpy as np
import pytest

from bsdf_cases import random_dirs


def _chords(rng, n, center, radius):
    """Uniformly random chords through a sphere (test_kd.cpp:98-113)."""
    a = random_dirs(rng, n) * radius + center
    b = random_dirs(rng, n) * radius + center
    d = b - a
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([a, np.full((n, 1), 1e-4), d, np.full((n, 1), np.inf)], 1).astype(np.float32)


@pytest.mark.parametrize("scene_name", ["cornell_box", "cornell_caustic"])
def test_kdtree_matches_bruteforce(pkg, oracle, scene_name):
    sb = getattr(pkg.scenes, scene_name)(64, 64)
    sc = oracle.scene(sb)
    rng = np.random.RandomState(0)
    rays = np.concatenate([_chords(rng, 40000, np.array([0, 1, 0], np.float32), 1.6),
                           sc.camera_rays((rng.rand(20000, 2) * 64).astype(np.float32))])
    tuv, prim, cnt = sc.trace(rays)
    tuv_b, prim_b = sc.trace_bruteforce(rays)
    assert np.array_equal(prim, prim_b)
    hit = prim != 0xFFFFFFFF
    assert hit.mean() > 0.5
    assert np.array_equal(tuv[hit], tuv_b[hit])  # same arithmetic, same primitive: bit-identical
    assert cnt["prims"] < 0.5 * len(rays) * sc.kd_info()["prims"]  # the tree actually culls


def test_kdtree_on_mesh(pkg, oracle):
    """A 2 x 80k-triangle heightfield exercises the min-max binning path (> exactPrimThreshold, gkdtree.h:1792-1925)."""
    S = pkg.scenes
    sb = S.SceneBuilder(32, 32)
    P, N, T = S.heightfield_mesh(n=201, seed=3)
    sb.trimesh(P, T, N=N, bsdf=sb.diffuse((0.5, 0.5, 0.5)))
    sb.rectangle([S.scale(0.2, 0.2, 1), S.rotate((1, 0, 0), 90), S.translate(0, 1, 0)], radiance=(5, 5, 5))
    sb.set_camera((0, 1.5, 2.5), (0, 0, 0), (0, 1, 0), 45.0)
    sc = oracle.scene(sb)
    info = sc.kd_info()
    assert info["prims"] == 2 * 200 * 200 + 1
    rng = np.random.RandomState(1)
    rays = _chords(rng, 3000, np.array([0, 0, 0], np.float32), 1.2)
    tuv, prim, _ = sc.trace(rays)
    tuv_b, prim_b = sc.trace_bruteforce(rays)
    assert np.array_equal(prim, prim_b)
    m = prim != 0xFFFFFFFF
    assert np.array_equal(tuv[m], tuv_b[m])
    # any-hit agrees with closest-hit on hit/miss
    _, occ, _ = sc.trace(rays, shadow=True)
    assert np.array_equal(occ != 0xFFFFFFFF, m)


def test_single_primitive_and_misses(pkg, oracle):
    S = pkg.scenes
    sb = S.SceneBuilder(8, 8)
    sb.rectangle([S.scale(1, 1, 1)], bsdf=sb.diffuse((0.5, 0.5, 0.5)), radiance=(1, 1, 1))
    sb.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0)
    sc = oracle.scene(sb)
    rays = np.array([[0, 0, 4, 1e-4, 0, 0, -1, np.inf], [3, 0, 4, 1e-4, 0, 0, -1, np.inf],
                     [0, 0, 4, 1e-4, 0, 0, 1, np.inf], [0.5, 0.5, 4, 1e-4, 0, 0, -1, 3.0]], np.float32)
    tuv, prim, _ = sc.trace(rays)
    assert prim.tolist() == [0, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF]
    assert tuv[0, 0] == 4.0 and tuv[0, 1] == 0.0 and tuv[0, 2] == 0.0  # rectangle returns local x, y


def _numpy_film(W, H, pos, rgb, stddev=0.5):
    """Direct restatement of ImageBlock::put on a single full-image block with border (no tiling)."""
    radius = 4 * stddev
    alpha = np.float32(-1.0 / (2 * stddev * stddev))
    xs = (np.float32(radius) * np.arange(31, dtype=np.float32)) / np.float32(31)
    vals = np.maximum(np.float32(0), np.exp(alpha * xs * xs) - np.exp(alpha * np.float32(radius) ** 2)).astype(np.float32)
    s = np.float32(0)
    for v in vals:
        s += v
    s *= np.float32(2 * radius / 31)
    vals = np.concatenate([vals * (np.float32(1) / s), [np.float32(0)]]).astype(np.float32)
    scale = np.float32(31 / radius)
    border = int(np.ceil(radius - 0.5))
    film = np.zeros((H, W, 5), np.float64)
    for (x, y), c in zip(pos, rgb):
        px, py = np.float32(x) - np.float32(0.5) + border, np.float32(y) - np.float32(0.5) + border
        x0, x1 = max(int(np.ceil(px - radius)), 0), min(int(np.floor(px + radius)), W + 2 * border - 1)
        y0, y1 = max(int(np.ceil(py - radius)), 0), min(int(np.floor(py + radius)), H + 2 * border - 1)
        for yy in range(y0, y1 + 1):
            wy = vals[min(int(abs((np.float32(yy) - py) * scale)), 31)]
            for xx in range(x0, x1 + 1):
                wx = vals[min(int(abs((np.float32(xx) - px) * scale)), 31)]
                fx, fy = xx - border, yy - border
                if 0 <= fx < W and 0 <= fy < H:
                    film[fy, fx] += np.float32(wx * wy) * np.array([c[0], c[1], c[2], 1, 1])
    return film, vals


def test_film_splat_against_numpy(pkg, oracle):
    sb = pkg.scenes.cornell_box(16, 12)  # smaller than one 32x32 tile: tiling does not enter
    sc = oracle.scene(sb)
    rng = np.random.RandomState(2)
    pos = (rng.rand(300, 2) * [16, 12]).astype(np.float32)
    rgb = rng.rand(300, 3).astype(np.float32)
    got = sc.film_splat(pos, rgb)
    want, vals = _numpy_film(16, 12, pos, rgb)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    # filter table: normalised so that its Riemann sum is 1 (rfilter.cpp:49-55)
    assert abs(vals[:31].sum() * 2 * 2.0 / 31 - 1.0) < 1e-6


def test_render_is_deterministic_and_partitionable(pkg, oracle):
    sb = pkg.scenes.cornell_box(64, 64)
    sc = oracle.scene(sb)
    p = pkg._abi.default_params()
    p.max_depth = 5
    a, sa = sc.render(p, 0, 2, nthreads=1)
    b, sb_ = sc.render(p, 0, 2, nthreads=4)
    for k in ("paths", "normal_rays", "shadow_rays", "path_length_sum"):
        assert sa[k] == sb_[k]
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5)  # only the merge order differs
    c, _ = sc.render(p, 0, 1)
    c, _ = sc.render(p, 1, 1, film=c)
    np.testing.assert_allclose(a, c, rtol=1e-5, atol=1e-5)
    assert sa["paths"] == 64 * 64 * 2 and 1.0 <= sa["path_length_sum"] / sa["paths"] <= 5.0


def test_triangle_clipping_known_answers(oracle):
    """test01_sutherlandHodgman of the reference (src/tests/test_kd.cpp:34-83): Triangle::getClippedAABB on the unit
    triangle -- the vectors the kd-tree's "perfect splits" rely on."""
    tri = [(0, 0, 0), (1, 0, 0), (1, 1, 0)]
    lo, hi = oracle.clipped_aabb(tri, (0, .5, -1), (1, 1, 1))          # split the triangle in half
    assert tuple(lo) == (.5, .5, 0) and tuple(hi) == (1, 1, 0)
    lo, hi = oracle.clipped_aabb(tri, (2, 2, 2), (3, 3, 3))            # completely clipped away
    assert not (lo <= hi).all()
    lo, hi = oracle.clipped_aabb(tri, (-1, -1, -1), (1, 1, 1))         # box contains the triangle: no clipping
    assert tuple(lo) == (0, 0, 0) and tuple(hi) == (1, 1, 0)
    lo, hi = oracle.clipped_aabb(tri, (-100, -100, 0), (100, 100, 0))  # triangle within a flat cell is kept
    assert tuple(lo) == (0, 0, 0) and tuple(hi) == (1, 1, 0)
    lo, hi = oracle.clipped_aabb(tri, (0, 1, 0), (1, 2, 0))            # just touching: collapsed point AABB
    assert tuple(lo) == (1, 1, 0) and tuple(hi) == (1, 1, 0)


# ---- src/tests/test_dgeom.cpp restated: known answers for ShapeKDTree::rayIntersect + fillIntersectionRecord -------------
def _one_triangle(pkg, oracle, N=None, UV=None):
    sb = pkg.scenes.SceneBuilder(8, 8, spp=1)
    sb.set_camera((0, 0, -4), (0, 0, 0), (0, 1, 0), 40.0)
    sb.trimesh(P=[[0, 0, 0], [1, 0, 0], [0, 1, 0]], T=[[0, 1, 2]], N=N, UV=UV, bsdf=sb.diffuse((0.5, 0.5, 0.5)))
    return oracle.scene(sb)


def test_dgeom_trimesh_known_answers(pkg, oracle):
    ray = np.array([[0.1, 0.2, -1.0, 0.0, 0, 0, 1, np.inf]], np.float32)   # Ray(Point(0.1, 0.2, -1), Vector(0, 0, 1))
    eps = 1e-4                                                             # Epsilon (constants.h:28)
    # test01_trimesh_1 (test_dgeom.cpp:34-66): no normals, no UV coordinates -- exact answers
    its = _one_triangle(pkg, oracle).intersect(ray)
    assert np.array_equal(its["p"][0], np.float32([0.1, 0.2, 0.0]))
    assert np.array_equal(its["uv"][0], np.float32([0.1, 0.2]))           # barycentric coordinates stand in for UVs
    assert np.array_equal(its["sh_n"][0], [0, 0, 1]) and np.array_equal(its["geo_n"][0], [0, 0, 1])
    assert np.array_equal(its["dpdu"][0], [1, 0, 0])
    assert its["t"][0] == 1.0
    # test02_trimesh_2 (:68-117): shading normals and UV coordinates, no explicit parameterisation
    normals = np.float32([[-0.3, 0, 1], [0.3, 0, 1], [0, 0.3, 1]])
    uv = np.float32([[0.1, 0.1], [1.1, 0.1], [0.1, 0.9]])
    its = _one_triangle(pkg, oracle, N=normals, UV=uv).intersect(ray)
    assert np.array_equal(its["p"][0], np.float32([0.1, 0.2, 0.0]))
    assert np.abs(its["uv"][0] - [0.2, 0.26]).max() <= eps
    assert np.array_equal(its["geo_n"][0], [0, 0, 1])
    n = normals[0] * 0.7 + normals[1] * 0.1 + normals[2] * 0.2
    n /= np.linalg.norm(n)
    assert np.abs(its["sh_n"][0] - n).max() <= eps
    assert np.abs(its["dpdu"][0] - [1, 0, 0]).max() <= eps                # vertices[1] - vertices[0]
    s = its["dpdu"][0] - its["sh_n"][0] * np.dot(its["sh_n"][0], its["dpdu"][0])
    assert np.abs(its["sh_s"][0] - s / np.linalg.norm(s)).max() <= eps    # computeShadingFrame (util.cpp:605-610)
    # a ray that misses the triangle
    miss = oracle_miss = _one_triangle(pkg, oracle).intersect(np.array([[0.8, 0.8, -1.0, 0.0, 0, 0, 1, np.inf]], np.float32))
    assert np.isinf(miss["t"][0])


# ---- next-event estimation, tested the way src/tests/test_chisquare.cpp test03_EmitterDirect tests an emitter ---------------
def test_emitter_direct_sampling_chi_square(pkg, oracle):
    """EmitterAdapter (test_chisquare.cpp:344-391): directions from sampleDirect against the solid-angle density pdfDirect on
    a 10 x 20 (theta, phi) grid, significance 0.0025 (Sidak over the reference points). The reference runs this on its
    envmap emitter only (data/tests/test_emitter.xml); here the same test drives the path's emitter, the area light
    (area.cpp:158-183, shape.cpp:102-126, scene.cpp:871-895, 992-995), at three unoccluded reference points of the Cornell box.
    Also: sampleDirect's value x pdf is the emitted radiance (17, 12, 4), the identity test_chisquare.cpp:34-38 checks for
    BSDFs."""
    from scipy import stats

    osc = oracle.scene(pkg.scenes.cornell_box(32, 32, spp=1))
    rng = np.random.RandomState(3)
    refs = ([-0.7, 0.2, 0.8], [0.0, 1.0, 0.9], [0.8, 1.5, 0.0])
    alpha = 1 - (1 - 0.0025) ** (1.0 / len(refs))
    theta_bins, phi_bins = 10, 20
    for ref in refs:
        zero_n = [0, 0, 0]                                  # DirectSamplingRecord(Point, time): no reference normal
        n = theta_bins * phi_bins * 1000
        d, dist, pdf_s, val = osc.emitter_sample(ref, zero_n, rng.rand(n, 2))
        assert (pdf_s > 0).all() and np.abs(np.linalg.norm(d, axis=1) - 1).max() < 1e-5
        np.testing.assert_allclose(val * pdf_s[:, None], np.tile([17.0, 12.0, 4.0], (n, 1)), rtol=1e-5)
        assert np.abs(osc.emitter_pdf(ref, zero_n, d[:5000]) - pdf_s[:5000]).max() <= 1e-5 * pdf_s.max()
        theta = np.arccos(np.clip(d[:, 2], -1, 1))
        phi = np.arctan2(d[:, 1], d[:, 0])
        phi[phi < 0] += 2 * np.pi
        ti = np.minimum((theta / np.pi * theta_bins).astype(int), theta_bins - 1)
        pi_ = np.minimum((phi / (2 * np.pi) * phi_bins).astype(int), phi_bins - 1)
        obs = np.bincount(ti * phi_bins + pi_, minlength=theta_bins * phi_bins).astype(np.float64)
        # expected bin masses: the integral of pdfDirect over each (theta, phi) cell. The density is discontinuous at the light's
        # silhouette, so instead of a quadrature over the sphere (the reference integrates adaptively, chisquare.cpp) the
        # integral is taken over the light's own area, where the integrand pdf(w(x)) |cos| / r^2 is smooth: x on a 1200 x 1200
        # midpoint grid of the 0.47 x 0.38 rectangle at y = 1.98 (scenes.cornell_box); only the cell boundaries cut the grid.
        m = 1200
        gx = ((np.arange(m) + 0.5) / m - 0.5) * 0.47
        gz = ((np.arange(m) + 0.5) / m - 0.5) * 0.38
        X, Z = np.meshgrid(gx, gz, indexing="ij")
        x = np.stack([X.ravel(), np.full(m * m, 1.98), Z.ravel()], 1)
        v = x - np.asarray(ref, np.float64)
        r2 = (v * v).sum(1)
        w_dir = (v / np.sqrt(r2)[:, None]).astype(np.float32)
        pdf = osc.emitter_pdf(ref, zero_n, w_dir).astype(np.float64)
        mass = pdf * np.abs(w_dir[:, 1]) / r2 * (0.47 * 0.38 / (m * m))       # the light's normal is -y
        assert abs(mass.sum() - 1) < 2e-3
        th = np.arccos(np.clip(w_dir[:, 2], -1, 1))
        ph = np.arctan2(w_dir[:, 1], w_dir[:, 0])
        ph[ph < 0] += 2 * np.pi
        cell = np.minimum((th / np.pi * theta_bins).astype(int), theta_bins - 1) * phi_bins + \
            np.minimum((ph / (2 * np.pi) * phi_bins).astype(int), phi_bins - 1)
        exp = np.bincount(cell, weights=mass, minlength=theta_bins * phi_bins) * n / mass.sum()
        chsq, dof, po, pe = 0.0, 0, 0.0, 0.0
        for i in np.argsort(exp):
            if exp[i] == 0:
                assert obs[i] <= n * 1e-3, "samples where the density is zero"
                continue
            if exp[i] < 5:
                po += obs[i]
                pe += exp[i]
                continue
            chsq += (obs[i] - exp[i]) ** 2 / exp[i]
            dof += 1
        if pe > 0:
            chsq += (po - pe) ** 2 / pe
            dof += 1
        p = float(stats.chi2.sf(chsq, max(dof - 1, 1)))
        assert p > alpha, "chi^2 rejected the area light at ref=%s (p=%g, dof=%d)" % (ref, p, dof)
    # facing tests of Scene::sampleEmitterDirect / AreaLight::sampleDirect: a reference normal pointing away from the light,
    # or a reference point behind the (one-sided) light, gets no contribution
    d, dist, pdf_s, val = osc.emitter_sample([0.0, 1.0, 0.9], [0, -1, 0], rng.rand(100, 2))
    assert (pdf_s == 0).all() and (val == 0).all()
    d, dist, pdf_s, val = osc.emitter_sample([0.0, 2.5, 0.0], [0, 0, 0], rng.rand(100, 2))
    assert (pdf_s == 0).all() and (val == 0).all()


@pytest.mark.parametrize("fov_axis", [0, 1, 2, 3, 4])
def test_perspective_camera_geometry(pkg, oracle, fov_axis):
    """PerspectiveCameraImpl (perspective.cpp:126-180, 271-298) and Transform::lookAt / perspective (transform.cpp:99-123,
    191-214), pinned to what they must satisfy: the centre ray runs along the viewing direction, the field of view is
    measured along the chosen axis (x, y, diagonal, smaller, larger: sensor.cpp:240-258), pixels are square, the image is
    not mirrored (+x to the right, pixel row 0 at the top for a camera on +z looking at the origin with up = +y), and the
    ray interval is the near / far clip distance measured along the optical axis."""
    W, H, fov = 64, 40, 50.0
    sb = pkg.scenes.SceneBuilder(W, H, spp=1)
    sb.rectangle([pkg.scenes.scale(1, 1, 1)], bsdf=sb.diffuse((0.5, 0.5, 0.5)), radiance=(1, 1, 1))
    origin, target = np.array([0.5, 0.25, 4.0]), np.array([0.0, 0.0, 0.0])
    sb.set_camera(tuple(origin), tuple(target), (0, 1, 0), fov, fov_axis=fov_axis)
    osc = oracle.scene(sb)
    pos = np.float32([[W / 2, H / 2], [0, H / 2], [W, H / 2], [W / 2, 0], [W / 2, H], [0, 0], [W, H]])
    r = osc.camera_rays(pos)
    o, mint, d, maxt = r[:, :3], r[:, 3], r[:, 4:7], r[:, 7]
    axis = (target - origin) / np.linalg.norm(target - origin)
    np.testing.assert_allclose(o, np.tile(origin, (len(pos), 1)), atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1, atol=1e-6)
    np.testing.assert_allclose(d[0], axis, atol=1e-6)
    cosang = d @ axis
    half = lambda i: np.degrees(np.arccos(cosang[i]))              # angle between ray i and the optical axis
    np.testing.assert_allclose(half(1), half(2), atol=1e-3)        # symmetric
    np.testing.assert_allclose(half(3), half(4), atol=1e-3)
    tx, ty = np.tan(np.radians(half(1))), np.tan(np.radians(half(3)))
    np.testing.assert_allclose(ty / tx, H / W, rtol=1e-4)          # square pixels
    tdiag = np.tan(np.radians(half(5)))
    np.testing.assert_allclose(tdiag, np.hypot(tx, ty), rtol=1e-4)
    measured = {0: 2 * half(1), 1: 2 * half(3), 2: 2 * half(5), 3: 2 * min(half(1), half(3)), 4: 2 * max(half(1), half(3))}[fov_axis]
    np.testing.assert_allclose(measured, fov, atol=2e-3)
    # orientation: pixel column 0 looks to -x, row 0 looks up
    right = np.cross(axis, [0, 1, 0])
    right /= np.linalg.norm(right)
    up = np.cross(right, axis)
    assert d[1] @ right < 0 < d[2] @ right and d[3] @ up > 0 > d[4] @ up
    # near / far clip (defaults 1e-2 / 1e4, sensor.cpp:158-160) along the optical axis: mint = near / cos, maxt = far / cos
    np.testing.assert_allclose(mint * cosang, 1e-2, rtol=1e-4)
    np.testing.assert_allclose(maxt * cosang, 1e4, rtol=1e-4)
