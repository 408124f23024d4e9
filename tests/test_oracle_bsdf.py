"""Pins the oracle's BSDF restatement the way the reference's own test does (src/tests/test_chisquare.cpp:393-620,
src/libcore/chisquare.cpp): (i) sample().weight * pdf == eval within ERROR_REQ = 1e-2 (test_chisquare.cpp:34-38),
(ii) chi^2 test of sample() against numerically integrated pdf() on a 10x20 (theta, phi) grid with
thetaBins*phiBins*1000 samples (chisquare.cpp:48-58) at significance 0.0025 with Sidak correction."""
import numpy as np
import pytest
from scipy import stats

from bsdf_cases import bsdf_scene, random_dirs

SMOOTH = ["diffuse", "twosided_diffuse", "roughconductor_beckmann_0.3", "roughconductor_ggx_0.15",
          "roughplastic_beckmann_0.7", "roughplastic_ggx_0.2_twosided"]
SIGNIFICANCE = 0.0025
N_WI = 6  # the reference uses 20 incident directions; 6 keeps the CPU suite short


@pytest.fixture(scope="module")
def scene(oracle):
    sb, idx = bsdf_scene()
    return oracle.scene(sb), idx


@pytest.mark.parametrize("name", SMOOTH)
def test_weight_times_pdf_equals_eval(scene, name):
    sc, idx = scene
    rng = np.random.RandomState(11)
    n = 20000
    wi = random_dirs(rng, n, upper=not name.count("twosided"))
    u = rng.rand(n, 2).astype(np.float32)
    s = sc.bsdf(idx[name], wi, wi, u)
    ok = s["spdf"] > 0
    assert ok.mean() > 0.5
    e = sc.bsdf(idx[name], wi[ok], s["wo"][ok], u[ok])
    lhs = s["weight"][ok] * s["spdf"][ok, None]
    rhs = e["eval"]
    # test_chisquare.cpp:530-560 compares per channel, relative for large values
    err = np.abs(lhs - rhs) / np.maximum(np.abs(rhs), 1.0)
    assert err.max() < 1e-2
    # the pdf returned by sample() and by pdf() must agree as well
    perr = np.abs(s["spdf"][ok] - e["pdf"]) / np.maximum(e["pdf"], 1.0)
    assert perr.max() < 1e-2


def _chi2(sc, index, wi, rng, theta_bins=10, phi_bins=20, sub=24):
    n = theta_bins * phi_bins * 1000
    u = rng.rand(n, 2).astype(np.float32)
    wis = np.repeat(wi[None], n, 0).astype(np.float32)
    s = sc.bsdf(index, wis, wis, u)
    ok = s["spdf"] > 0
    wo = s["wo"][ok]
    theta = np.arccos(np.clip(wo[:, 2], -1, 1))
    phi = np.arctan2(wo[:, 1], wo[:, 0])
    phi[phi < 0] += 2 * np.pi
    ti = np.minimum((theta / np.pi * theta_bins).astype(int), theta_bins - 1)
    pi_ = np.minimum((phi / (2 * np.pi) * phi_bins).astype(int), phi_bins - 1)
    obs = np.bincount(ti * phi_bins + pi_, minlength=theta_bins * phi_bins).astype(np.float64)
    # expected frequencies: midpoint quadrature of pdf * sin(theta) over each bin
    tt = (np.arange(theta_bins * sub) + 0.5) * (np.pi / (theta_bins * sub))
    pp = (np.arange(phi_bins * sub) + 0.5) * (2 * np.pi / (phi_bins * sub))
    T, P = np.meshgrid(tt, pp, indexing="ij")
    d = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3).astype(np.float32)
    pdf = sc.bsdf(index, np.repeat(wi[None], d.shape[0], 0).astype(np.float32), d, np.zeros((d.shape[0], 2), np.float32))["pdf"]
    w = (pdf.reshape(T.shape) * np.sin(T)) * (np.pi / (theta_bins * sub)) * (2 * np.pi / (phi_bins * sub))
    exp = w.reshape(theta_bins, sub, phi_bins, sub).sum((1, 3)).ravel() * n
    # pool low-frequency cells (chisquare.cpp: minExpFrequency = 5)
    order = np.argsort(exp)
    pooled_o = pooled_e = 0.0
    chsq, dof = 0.0, 0
    for i in order:
        if exp[i] == 0:
            if obs[i] > n * 1e-5:
                return 0.0  # samples in a region of zero density
            continue
        if exp[i] < 5:
            pooled_o += obs[i]
            pooled_e += exp[i]
            continue
        chsq += (obs[i] - exp[i]) ** 2 / exp[i]
        dof += 1
    if pooled_e > 0:
        chsq += (pooled_o - pooled_e) ** 2 / pooled_e
        dof += 1
    return float(stats.chi2.sf(chsq, max(dof - 1, 1)))


@pytest.mark.parametrize("name", SMOOTH)
def test_chi_square_sampling_vs_pdf(scene, name):
    sc, idx = scene
    rng = np.random.RandomState(5)
    alpha = 1 - (1 - SIGNIFICANCE) ** (1.0 / N_WI)  # Sidak, test_chisquare.cpp
    for k in range(N_WI):
        wi = random_dirs(rng, 1, upper=True)[0]
        wi[2] = max(wi[2], 0.05)
        wi /= np.linalg.norm(wi)
        p = _chi2(sc, idx[name], wi, rng)
        assert p > alpha, "chi^2 rejected %s at wi=%s (p=%g)" % (name, wi, p)


def test_dielectric_delta_lobes(scene):
    """dielectric.cpp:289-340: reflection with probability F, refraction otherwise; weights 1 and eta^-2-type factors."""
    sc, idx = scene
    rng = np.random.RandomState(3)
    n = 50000
    wi = random_dirs(rng, n)
    u = rng.rand(n, 2).astype(np.float32)
    s = sc.bsdf(idx["dielectric_water_air"], wi, wi, u)
    refl = (s["flags"] & 0x20) != 0
    trans = (s["flags"] & 0x40) != 0
    assert (refl ^ trans).all()
    # reflection mirrors wi about the normal, weight = specularReflectance = 1
    np.testing.assert_allclose(s["wo"][refl], wi[refl] * np.array([-1, -1, 1], np.float32), atol=1e-6)
    np.testing.assert_allclose(s["weight"][refl], 1.0, atol=1e-6)
    # refraction: Snell's law and the radiance scaling factor (eta_i/eta_t)^2
    eta = np.float32(1.3330) / np.float32(1.000277)
    sin_i = np.sqrt(1 - wi[trans, 2] ** 2)
    sin_t = np.sqrt(1 - s["wo"][trans, 2] ** 2)
    ratio = np.where(wi[trans, 2] > 0, 1 / eta, eta)
    np.testing.assert_allclose(sin_t, sin_i * ratio, atol=2e-5)
    np.testing.assert_allclose(s["weight"][trans, 0], ratio ** 2, rtol=1e-5)
    assert (np.sign(s["wo"][trans, 2]) == -np.sign(wi[trans, 2])).all()
    # probability of reflection equals the Fresnel term: check against the empirical frequency at normal incidence
    wi0 = np.repeat(np.array([[0, 0, 1]], np.float32), n, 0)
    s0 = sc.bsdf(idx["dielectric_water_air"], wi0, wi0, u)
    F = ((eta - 1) / (eta + 1)) ** 2
    assert abs(((s0["flags"] & 0x20) != 0).mean() - F) < 3 * np.sqrt(F / n) + 1e-4
    # eval/pdf w.r.t. solid angle are zero for delta lobes
    e = sc.bsdf(idx["dielectric_water_air"], wi, s["wo"], u)
    assert (e["eval"] == 0).all() and (e["pdf"] == 0).all()


def test_energy_conservation(scene):
    """E[weight] <= 1 for every model with unit specular reflectance (albedo of a BRDF cannot exceed one)."""
    sc, idx = scene
    rng = np.random.RandomState(9)
    n = 200000
    for name in SMOOTH:
        wi = np.repeat(np.array([[0.3, 0.1, 0.9486833]], np.float32), n, 0)
        s = sc.bsdf(idx[name], wi, wi, rng.rand(n, 2).astype(np.float32))
        assert s["weight"].mean(0).max() <= 1.0 + 5e-3, name


# ---- the reference's rough-transmittance test (src/tests/test_rtrans.cpp:135-260): the tabulated transmittance, reduced
# to one (eta, alpha) like RoughPlastic::configure does (rtrans.h:292-388), against a numerical integration of the rough
# dielectric BTDF (tolerance 1e-3 there; a few 1e-3 here because the integral is Monte Carlo) ----------------------------
def _cubic_interp(x, values):  # evalCubicInterp1D on [0, 1] (src/libcore/spline.cpp:23-60)
    size = len(values)
    t = x * (size - 1)
    k = max(0, min(int(t), size - 2))
    f0, f1 = values[k], values[k + 1]
    d0 = 0.5 * (values[k + 1] - values[k - 1]) if k > 0 else values[k + 1] - values[k]
    d1 = 0.5 * (values[k + 2] - values[k]) if k + 2 < size else values[k + 1] - values[k]
    t -= k
    t2, t3 = t * t, t * t * t
    return (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1


def _fresnel_dielectric(c, eta):  # fresnelDielectricExt for cos >= 0, eta > 1 (util.cpp:653-683)
    st2 = (1 - c * c) / (eta * eta)
    ct = np.sqrt(np.maximum(1 - st2, 0))
    rs = (c - eta * ct) / (c + eta * ct)
    rp = (eta * c - ct) / (eta * c + ct)
    return 0.5 * (rs * rs + rp * rp), ct


def _g1(v, m, alpha, ggx):  # smithG1 (microfacet.h:477-514)
    ok = (np.einsum("ij,ij->i", v, m) * v[:, 2]) > 0
    tan = np.sqrt(np.maximum(1 - v[:, 2] ** 2, 0)) / np.abs(v[:, 2])
    if ggx:
        g = 2 / (1 + np.sqrt(1 + (alpha * tan) ** 2))
    else:
        a = 1 / np.maximum(alpha * tan, 1e-12)
        g = np.where(a >= 1.6, 1.0, (3.535 * a + 2.181 * a * a) / (1 + 2.276 * a + 2.577 * a * a))
    return np.where(tan == 0, 1.0, g) * ok


@pytest.mark.parametrize("distr,eta,alpha", [("beckmann", 1.5, 0.2), ("ggx", 1.3, 0.4), ("beckmann", 1.49, 0.7)])
def test_rough_transmittance_table_vs_btdf_integration(pkg, oracle, distr, eta, alpha):
    from b200pg import rtrans

    ext, ed, idf = oracle.rtrans_reduce(rtrans.load_packed(distr), eta, alpha)
    rng = np.random.RandomState(4)
    n = 400000
    ggx = distr == "ggx"
    for cos_i in (0.25, 0.5, 0.8, 1.0):
        # microfacet normals m ~ D(m) cos(theta_m) (sampleAll, microfacet.h:287-395)
        u1, u2 = rng.rand(n), rng.rand(n)
        tan2 = alpha * alpha * u1 / (1 - u1) if ggx else -alpha * alpha * np.log(1 - u1)
        cm = 1 / np.sqrt(1 + tan2)
        sm = np.sqrt(np.maximum(1 - cm * cm, 0))
        phi = 2 * np.pi * u2
        m = np.stack([sm * np.cos(phi), sm * np.sin(phi), cm], 1)
        wi = np.array([np.sqrt(1 - cos_i * cos_i), 0.0, cos_i])
        wim = m @ wi
        F, ct = _fresnel_dielectric(np.maximum(wim, 0), eta)
        # refracted direction (util.cpp refract): wo = m * (wi.m / eta - cos_t) - wi / eta
        wo = m * (wim / eta - ct)[:, None] - wi[None] / eta
        wiv = np.repeat(wi[None], n, 0)
        # weight of the transmission lobe for pdf D cos_m: |wi.m| G / (cos_i cos_m) * (1 - F)   (Walter et al. 2007, eq. 41)
        w = np.where(wim > 0, (1 - F) * wim * _g1(wiv, m, alpha, ggx) * _g1(wo, m, alpha, ggx) / (cos_i * cm), 0.0)
        ref = w.mean()
        dat = _cubic_interp(cos_i ** 0.25, ext)  # RoughTransmittance::eval warps cos(theta) with pow(., 1/4), rtrans.h:232-240
        assert abs(ref - dat) < 4e-3 + 4 * w.std() / np.sqrt(n), (distr, eta, alpha, cos_i, ref, dat)
    # diffuse transmittance = 2 int_0^1 x T(x) dx (test_rtrans.cpp:44-47) of the same table
    xs = (np.arange(4000) + 0.5) / 4000
    num = np.mean([2 * x * _cubic_interp(x ** 0.25, ext) for x in xs])
    assert abs(num - ed) < 3e-3, (num, ed)


# ---- src/tests/test_microfacet.cpp restated (the reference's own test of MicrofacetDistribution) ---------------------------
def _chi2_visible(oracle, distr, au, av, wi, rng, theta_bins=10, phi_bins=20, sub=24):
    """ChiSquare(thetaBins = 10, 2 * thetaBins) of sampleVisible(wi, .) against pdfVisible(wi, .) (test_microfacet.cpp:133-170)."""
    n = theta_bins * phi_bins * 1000
    m, pdf_s, _, _ = oracle.microfacet(distr, au, av, wi, u=rng.rand(n, 2))
    # the adapter's assertions (test_microfacet.cpp:72-76): finite, unit length
    assert np.isfinite(m).all() and np.abs(np.linalg.norm(m, axis=1) - 1).max() < 1e-4
    assert np.isfinite(pdf_s).all() and (pdf_s >= 0).all()
    theta = np.arccos(np.clip(m[:, 2], -1, 1))
    phi = np.arctan2(m[:, 1], m[:, 0])
    phi[phi < 0] += 2 * np.pi
    ti = np.minimum((theta / np.pi * theta_bins).astype(int), theta_bins - 1)
    pi_ = np.minimum((phi / (2 * np.pi) * phi_bins).astype(int), phi_bins - 1)
    obs = np.bincount(ti * phi_bins + pi_, minlength=theta_bins * phi_bins).astype(np.float64)
    tt = (np.arange(theta_bins * sub) + 0.5) * (np.pi / (theta_bins * sub))
    pp = (np.arange(phi_bins * sub) + 0.5) * (2 * np.pi / (phi_bins * sub))
    T, P = np.meshgrid(tt, pp, indexing="ij")
    d = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3).astype(np.float32)
    _, pdf, _, _ = oracle.microfacet(distr, au, av, wi, m=d)
    w = (pdf.astype(np.float64).reshape(T.shape) * np.sin(T)) * (np.pi / (theta_bins * sub)) * (2 * np.pi / (phi_bins * sub))
    total = w.sum()
    exp = w.reshape(theta_bins, sub, phi_bins, sub).sum((1, 3)).ravel() * n
    order = np.argsort(exp)
    pooled_o = pooled_e = 0.0
    chsq, dof = 0.0, 0
    for i in order:  # chisquare.cpp: cells with an expected frequency below 5 are pooled
        if exp[i] == 0:
            if obs[i] > n * 1e-5:
                return 0.0, total
            continue
        if exp[i] < 5:
            pooled_o += obs[i]
            pooled_e += exp[i]
            continue
        chsq += (obs[i] - exp[i]) ** 2 / exp[i]
        dof += 1
    if pooled_e > 0:
        chsq += (pooled_o - pooled_e) ** 2 / pooled_e
        dof += 1
    return float(stats.chi2.sf(chsq, max(dof - 1, 1))), total


def test_microfacet_visible_normal_sampling(oracle):
    """test02_MicrofacetVisible (test_microfacet.cpp:133-170): Beckmann 0.3, Beckmann 0.5/0.3, GGX 0.1, GGX 0.2/0.3, each at 10
    incident directions drawn uniformly from the hemisphere; significance 0.0025 with Sidak correction over the 40 tests."""
    rng = np.random.RandomState(11)
    distrs = [(0, 0.3, 0.3), (0, 0.5, 0.3), (1, 0.1, 0.1), (1, 0.2, 0.3)]
    n_tests = 10 * len(distrs)
    alpha = 1 - (1 - SIGNIFICANCE) ** (1.0 / n_tests)
    for k in range(10):
        u = rng.rand(2)
        z = u[0]                                                   # warp::squareToUniformHemisphere (warp.cpp:33-39)
        r = np.sqrt(max(0.0, 1 - z * z))
        wi = np.array([r * np.cos(2 * np.pi * u[1]), r * np.sin(2 * np.pi * u[1]), z], np.float32)
        for distr, au, av in distrs:
            p, total = _chi2_visible(oracle, distr, au, av, wi, rng)
            assert p > alpha, "chi^2 rejected distr %d alpha (%g, %g) at wi=%s (p=%g)" % (distr, au, av, wi, p)
            if wi[2] > 0.2:  # the visible-normal density integrates to one (midpoint rule on the test's own grid)
                assert abs(total - 1) < 2e-2, (distr, au, av, wi, total)


@pytest.mark.parametrize("distr,au,av", [(0, 0.5, 0.5), (0, 0.5, 0.3), (1, 0.5, 0.5), (1, 0.5, 0.3)])
def test_microfacet_projected_area_normalisation(oracle, distr, au, av):
    """pdfAll(m) = D(m) cos(theta_m) is a density (what test01_Microfacet checks for sampleAll, test_microfacet.cpp:95-131, on
    the same four Beckmann / GGX parameter sets): the projected microfacet area integrates to one; and Smith's G1 at normal
    incidence is one."""
    nt, npi = 2000, 720
    tt = (np.arange(nt) + 0.5) * (0.5 * np.pi / nt)
    pp = (np.arange(npi) + 0.5) * (2 * np.pi / npi)
    T, P = np.meshgrid(tt, pp, indexing="ij")
    m = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3).astype(np.float32)
    _, _, D, G1 = oracle.microfacet(distr, au, av, np.array([0, 0, 1], np.float32), m=m)
    integral = (D.astype(np.float64).reshape(T.shape) * np.cos(T) * np.sin(T)).sum() * (0.5 * np.pi / nt) * (2 * np.pi / npi)
    assert abs(integral - 1) < 2e-3, integral
    assert np.all(G1[m[:, 2] > 0] == 1.0)
