"""Pins the oracle's medium restatement without a GPU.

The reference's own test for this part is test02_PhaseFunction of src/tests/test_chisquare.cpp:510-575 on
data/tests/test_phase.xml (isotropic, hg g = 0.9, hg g = -0.3): chi^2 of sample() against pdf() plus the identity
sample-weight == eval/pdf (which is 1 for both models). There is no reference test for the heterogeneous medium or
the grid volume; those are pinned to what the algorithm must satisfy analytically:
  * trilinear lookup reproduces a linear field exactly and is zero outside (gridvolume.cpp:337-388)
  * Woodcock tracking in a constant-density grid samples an exponential free-flight distance with rate
    sigma_t = scale * density, and the ratio-free transmittance estimator has mean exp(-sigma_t * length)
    (heterogeneous.cpp:546-663)
"""
import numpy as np
import pytest
from scipy import stats

from bsdf_cases import random_dirs


def _medium_scene(pkg, dens, phase="hg", g=0.7, scale_=5.0, lo=(-1, -1, -1), hi=(1, 1, 1)):
    S = pkg.scenes
    sb = S.SceneBuilder(16, 16)
    med = sb.medium(dens, lo, hi, scale_=scale_, albedo=(0.9, 0.9, 0.9), phase=phase, g=g)
    sb.cube([S.scale(1, 1, 1)], bsdf=-1, interior=med)
    sb.rectangle([S.scale(0.2, 0.2, 1), S.rotate((1, 0, 0), 90), S.translate(0, 3, 0)], radiance=(5, 5, 5))
    sb.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0)
    return sb


def test_grid_lookup_is_trilinear(pkg, oracle):
    res = 9
    z, y, x = np.meshgrid(np.linspace(0, 1, res), np.linspace(0, 1, res), np.linspace(0, 1, res), indexing="ij")
    dens = (0.2 + 0.3 * x + 0.1 * y + 0.25 * z).astype(np.float32)  # linear in the grid coordinates, <= 1
    sc = oracle.scene(_medium_scene(pkg, dens))
    rng = np.random.RandomState(0)
    P = (rng.rand(20000, 3) * 2 - 1).astype(np.float32) * 0.999
    want = 0.2 + 0.3 * (P[:, 0] + 1) / 2 + 0.1 * (P[:, 1] + 1) / 2 + 0.25 * (P[:, 2] + 1) / 2
    got = sc.grid_lookup(0, P)
    np.testing.assert_allclose(got, want, atol=2e-6)
    # exactly on grid nodes: the stored values (x2 = x1 + 1 must exist, so the last layer is "outside", :353-356)
    nodes = np.stack([x, y, z], -1)[:-1, :-1, :-1].reshape(-1, 3).astype(np.float32) * 2 - 1
    np.testing.assert_allclose(sc.grid_lookup(0, nodes), dens[:-1, :-1, :-1].ravel(), atol=1e-6)
    outside = np.array([[1.01, 0, 0], [0, -1.5, 0], [0, 0, 7], [-1.0001, 0, 0]], np.float32)
    assert (sc.grid_lookup(0, outside) == 0).all()


def _rays(o, d, mint=0.0, maxt=np.inf):
    n = o.shape[0]
    return np.concatenate([o, np.full((n, 1), mint), d, np.full((n, 1), maxt)], 1).astype(np.float32)


@pytest.mark.parametrize("density,scale_", [(1.0, 4.0), (0.35, 7.0)])
def test_woodcock_free_flight_and_transmittance(pkg, oracle, density, scale_):
    dens = np.full((8, 8, 8), density, np.float32)
    sc = oracle.scene(_medium_scene(pkg, dens, scale_=scale_))
    sigma = scale_ * density
    n = 200000
    # rays along +x through the middle of the box, entering at x = -1; the grid's last cell layer is empty
    # (x2 >= res -> 0), so keep the segment inside [-1, 0.7]
    o = np.repeat(np.array([[-1.0, 0.05, -0.1]], np.float32), n, 0)
    d = np.repeat(np.array([[1.0, 0, 0]], np.float32), n, 0)
    length = 1.5
    t, tr, wo, pdf = sc.medium_sample(0, _rays(o, d, 0.0, length))
    scattered = np.isfinite(t)
    # P(scatter before length) = 1 - exp(-sigma * length)
    p = 1 - np.exp(-sigma * length)
    assert abs(scattered.mean() - p) < 4 * np.sqrt(p * (1 - p) / n)
    # conditional distances follow the truncated exponential: Kolmogorov-Smirnov
    cdf = lambda x: (1 - np.exp(-sigma * x)) / p
    ks = stats.kstest(t[scattered][:50000], cdf)
    assert ks.pvalue > 1e-3, ks
    # transmittance estimator: 2 ratio-free trials -> values in {0, .5, 1}, mean exp(-sigma * length)
    assert set(np.unique(tr)).issubset({0.0, 0.5, 1.0})
    T = np.exp(-sigma * length)
    assert abs(tr.mean() - T) < 4 * np.sqrt(T * (1 - T) / (2 * n)) + 1e-4
    # rays that miss the density box: no event, transmittance 1
    o2 = o + np.array([0, 5, 0], np.float32)
    t2, tr2, _, _ = sc.medium_sample(0, _rays(o2, d, 0.0, length))
    assert np.isinf(t2).all() and (tr2 == 1).all()


@pytest.mark.parametrize("phase,g", [("isotropic", 0.0), ("hg", 0.9), ("hg", -0.3), ("hg", 0.0)])
def test_phase_function_chi_square(pkg, oracle, phase, g):
    """test_chisquare.cpp:510-575 on the models of data/tests/test_phase.xml:5-21."""
    sc = oracle.scene(_medium_scene(pkg, np.full((4, 4, 4), 0.5, np.float32), phase=phase, g=g))
    rng = np.random.RandomState(2)
    theta_bins, phi_bins, sub = 10, 20, 16
    n = theta_bins * phi_bins * 1000
    n_wi = 4
    alpha = 1 - (1 - 0.0025) ** (1.0 / n_wi)
    for wi in random_dirs(rng, n_wi):
        wis = np.repeat(wi[None], n, 0)
        u = rng.rand(n, 2).astype(np.float32)
        ev, wo, pdf = sc.phase(0, wis, wis, u)
        # sample(): weight is 1, i.e. the returned pdf equals eval() of the sampled direction (hg.cpp:93-95)
        ev2, _, _ = sc.phase(0, wis[:20000], wo[:20000], u[:20000])
        np.testing.assert_allclose(ev2, pdf[:20000], rtol=1e-4)
        np.testing.assert_allclose(np.linalg.norm(wo, axis=1), 1.0, atol=1e-5)
        theta = np.arccos(np.clip(wo[:, 2], -1, 1))
        phi = np.arctan2(wo[:, 1], wo[:, 0])
        phi[phi < 0] += 2 * np.pi
        ti = np.minimum((theta / np.pi * theta_bins).astype(int), theta_bins - 1)
        pj = np.minimum((phi / (2 * np.pi) * phi_bins).astype(int), phi_bins - 1)
        obs = np.bincount(ti * phi_bins + pj, minlength=theta_bins * phi_bins).astype(np.float64)
        tt = (np.arange(theta_bins * sub) + 0.5) * (np.pi / (theta_bins * sub))
        pp = (np.arange(phi_bins * sub) + 0.5) * (2 * np.pi / (phi_bins * sub))
        T, P = np.meshgrid(tt, pp, indexing="ij")
        dirs = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3).astype(np.float32)
        pd, _, _ = sc.phase(0, np.repeat(wi[None], dirs.shape[0], 0), dirs, np.zeros((dirs.shape[0], 2), np.float32))
        w = pd.reshape(T.shape) * np.sin(T) * (np.pi / (theta_bins * sub)) * (2 * np.pi / (phi_bins * sub))
        assert abs(w.sum() - 1) < 2e-3  # the pdf integrates to one over the sphere
        exp = w.reshape(theta_bins, sub, phi_bins, sub).sum((1, 3)).ravel() * n
        big = exp >= 5
        chsq = ((obs[big] - exp[big]) ** 2 / exp[big]).sum()
        dof = int(big.sum())
        if (~big).any():
            chsq += (obs[~big].sum() - exp[~big].sum()) ** 2 / max(exp[~big].sum(), 1e-9)
            dof += 1
        assert stats.chi2.sf(chsq, dof - 1) > alpha, (phase, g, wi)
    # mean cosine of HG equals g (PhaseFunction::getMeanCosine, hg.cpp:112-114), with wi = incident direction
    # pointing away: cos between -wi and wo
    wi = np.array([0, 0, 1], np.float32)
    _, wo, _ = sc.phase(0, np.repeat(wi[None], 200000, 0), np.repeat(wi[None], 200000, 0), rng.rand(200000, 2).astype(np.float32))
    assert abs((-wo[:, 2]).mean() - g) < 5e-3


def test_volumetric_render_is_energy_consistent(pkg, oracle):
    """A purely absorbing (albedo 0) constant medium between camera and a diffuse emitter attenuates the directly
    visible emitter by exp(-sigma * thickness): ties Li_volpath, sampleDistance and the null-boundary handling together."""
    from b200pg import api

    S = pkg.scenes
    sb = S.SceneBuilder(16, 16, spp=1)
    dens = np.full((6, 6, 6), 1.0, np.float32)
    # the box spans z in [-0.5, 0.5]; the density grid is made larger so that the whole thickness is inside valid cells
    med = sb.medium(dens, (-2, -2, -0.75), (2, 2, 1.0), scale_=1.3, albedo=(0.0, 0.0, 0.0), phase="isotropic", g=0.0)
    sb.cube([S.scale(1.5, 1.5, 0.5)], bsdf=-1, interior=med)
    sb.rectangle([S.scale(4, 4, 1), S.translate(0, 0, -2)], radiance=(1, 1, 1))  # big emitter facing +z behind the slab
    sb.set_camera((0, 0, 6), (0, 0, 0), (0, 1, 0), 5.0)
    sc = oracle.scene(sb)
    p = api.default_params()
    p.max_depth = 8
    p.volumetric = 1
    rng = np.random.RandomState(1)
    n = 100000
    pix = rng.randint(0, 256, n).astype(np.uint32)
    smp = np.arange(n).astype(np.uint32)
    L = sc.radiance(p, pix, smp)
    T = np.exp(-1.3 * 1.0)
    assert set(np.unique(L[:, 0])).issubset({0.0, 1.0})  # the camera ray either reaches the emitter or is absorbed
    assert abs(L[:, 0].mean() - T) < 4 * np.sqrt(T * (1 - T) / n)


def test_guided_volumetric_path_is_unbiased(pkg, oracle):
    """Guided direction sampling at medium / surface vertices and guided collision probabilities (weighted delta
    tracking) change the sampling, not the expectation: the guided mean image matches the unguided one."""
    from b200pg import api
    from oracle_lib import develop

    sb = pkg.scenes.cornell_medium(40, 40, spp=8, res=24, scale_=12.0)
    sc = oracle.scene(sb)
    p = api.default_params()
    p.max_depth = 6
    p.volumetric = 1
    p.guiding_probability = 0.5
    fld = oracle.field(8, (-1.1, -0.1, -1.1), (1.1, 2.1, 1.1))
    sink = oracle.samples()
    for it in range(4):
        sink.clear()
        sc.render(p, 100 * it, 8, field=fld if it else None, sink=sink)
        s = sink.get()
        assert len(s["weight"]) > 10000 and np.isfinite(s["weight"]).all() and (s["weight"] >= 0).all()
        assert np.allclose(np.linalg.norm(s["dir"], axis=1), 1, atol=1e-4)
        # medium vertices are recorded too: positions strictly inside the medium box, away from every wall
        inside = (np.abs(s["pos"][:, 0]) < 0.59) & (np.abs(s["pos"][:, 2]) < 0.59) & (s["pos"][:, 1] > 0.21) & (s["pos"][:, 1] < 1.39)
        assert inside.mean() > 0.05
        fld.train(s, 4, 4000)
    assert fld.info()["cells"] > 1
    ref = develop(sc.render(p, 5000, 768)[0])
    tol = 0.02 * ref.mean()
    g_dir = develop(sc.render(p, 0, 384, field=fld)[0])
    assert abs(g_dir.mean() - ref.mean()) < tol
    p.guided_distance = 1
    g_dist = develop(sc.render(p, 0, 384, field=fld)[0])
    assert abs(g_dist.mean() - ref.mean()) < tol
    # region-wise too (the medium cube covers the image centre): guards against compensating errors
    c = slice(12, 28)
    assert abs(g_dist[c, c].mean() - ref[c, c].mean()) < 0.03 * ref[c, c].mean()
    assert abs(g_dir[c, c].mean() - ref[c, c].mean()) < 0.03 * ref[c, c].mean()


def test_simpson_quadrature_and_inversion(pkg, oracle):
    """method = simpson (heterogeneous.cpp:301-376, 420-544): composite Simpson is exact for a density that is linear along
    the ray, so both the transmittance exp(-int) and the inverted free-flight distance have closed forms."""
    import b200pg._abi as A

    res = 33
    z, y, x = np.meshgrid(np.linspace(0, 1, res), np.linspace(0, 1, res), np.linspace(0, 1, res), indexing="ij")
    dens = (0.2 + 0.6 * x).astype(np.float32)  # linear in world x over [-1, 1]: d(X) = 0.5 + 0.3 X
    sb = _medium_scene(pkg, dens, scale_=3.0)
    sb.media[0]["method"] = A.MEDIUM_SIMPSON
    sc = oracle.scene(sb)
    n = 100000
    o = np.repeat(np.array([[-0.9, 0.1, -0.2]], np.float32), n, 0)
    d = np.repeat(np.array([[1.0, 0, 0]], np.float32), n, 0)
    length = 1.6  # stays inside the valid cells (the last cell layer of a grid volume is empty)
    t, tr, wo, pdf = sc.medium_sample(0, _rays(o, d, 0.0, length))
    sigma = lambda X: 3.0 * (0.5 + 0.3 * X)
    integral = lambda s: 3.0 * ((0.5 + 0.3 * -0.9) * s + 0.15 * s * s)  # int_0^s sigma(-0.9 + u) du
    T = np.exp(-integral(length))
    # transmittance is deterministic with this method
    np.testing.assert_allclose(tr, T, rtol=2e-4)
    scattered = np.isfinite(t)
    assert abs(scattered.mean() - (1 - T)) < 4 * np.sqrt(T * (1 - T) / n)
    # the sampled distances follow the cdf 1 - exp(-integral(s)), truncated at `length`
    ks = stats.kstest(t[scattered][:50000], lambda s: (1 - np.exp(-integral(s))) / (1 - T))
    assert ks.pvalue > 1e-3, ks
    # the whole volumetric Li agrees between the two methods in expectation (absorbing slab, emitter behind it)
    from b200pg import api

    S = pkg.scenes
    means = []
    for method in (A.MEDIUM_WOODCOCK, A.MEDIUM_SIMPSON):
        sb2 = S.SceneBuilder(16, 16, spp=1)
        med = sb2.medium(np.full((6, 6, 6), 0.7, np.float32), (-2, -2, -0.75), (2, 2, 1.0), scale_=1.3, albedo=(0.5, 0.5, 0.5),
                         phase="isotropic", g=0.0)
        sb2.media[0]["method"] = method
        sb2.cube([S.scale(1.5, 1.5, 0.5)], bsdf=-1, interior=med)
        sb2.rectangle([S.scale(4, 4, 1), S.translate(0, 0, -2)], radiance=(1, 1, 1))
        sb2.set_camera((0, 0, 6), (0, 0, 0), (0, 1, 0), 5.0)
        p = api.default_params()
        p.max_depth, p.volumetric = 6, 1
        rng = np.random.RandomState(1)
        m = 60000
        L = oracle.scene(sb2).radiance(p, rng.randint(0, 256, m).astype(np.uint32), np.arange(m).astype(np.uint32))
        means.append(L[:, 0].mean())
    assert abs(means[0] - means[1]) < 0.02 * means[0], means
