"""Analytic pins of the oracle's light transport (ProgressiveMIPathTracer::Li restated, progressive_path.cpp:133-314): the
reference ships no rendered image or radiance value to compare with, so the whole estimator -- emitter
sampling, BSDF sampling, the power-heuristic MIS between them, Russian roulette -- is held against closed forms:
  * direct illumination of a diffuse floor point under a rectangular Lambertian light = rho * L * F, F the form factor of a
    differential element to a parallel rectangle;
  * the furnace: inside a closed box whose walls all emit L and reflect rho, the radiance is L / (1 - rho) in every direction.
The GPU path agrees with the oracle sample by sample (tests/test_gpu_parity.py), so these pins carry over."""
import numpy as np
import pytest

from transport_cases import corner_form_factor as _corner_form_factor, form_factor_scene, furnace_scene


def _params(pkg, **kw):
    p = pkg._abi.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("light", ["rectangle", "trimesh"])
@pytest.mark.parametrize("use_nee", [1, 0])
def test_direct_illumination_matches_the_form_factor(pkg, oracle, use_nee, light):
    sb, centre, want = form_factor_scene(pkg, light)
    osc = oracle.scene(sb)
    n = 400000
    pix = np.full(n, centre, np.uint32)
    rad = osc.radiance(_params(pkg, max_depth=2, use_nee=use_nee), pix, np.arange(n, dtype=np.uint32)).astype(np.float64)
    mean, sem = rad.mean(0), rad.std(0) / np.sqrt(n)
    assert np.all(np.abs(mean - want) <= 4 * sem + 2e-3 * want), (mean, want, sem)
    assert np.all(sem < 0.01 * want)


@pytest.mark.parametrize("kw", [dict(max_depth=-1, rr_depth=5), dict(max_depth=-1, rr_depth=1), dict(max_depth=-1, rr_depth=3, use_nee=0)])
def test_furnace(pkg, oracle, kw):
    sb, want = furnace_scene(pkg, res=8)
    rho, L = 0.5, 1.0
    osc = oracle.scene(sb)
    # every face really faces inwards: a ray from the centre sees an emitting front side everywhere
    rng = np.random.RandomState(0)
    n = 200000
    pix = rng.randint(0, 64, n).astype(np.uint32)
    rad = osc.radiance(_params(pkg, **kw), pix, np.arange(n, dtype=np.uint32)).astype(np.float64)
    mean, sem = rad.mean(0), rad.std(0) / np.sqrt(n)
    assert np.all(np.abs(mean - want) <= 4 * sem + 1e-3 * want), (mean, want, sem)
    # truncated series: maxDepth = k keeps the first k terms L (1 + rho + ... + rho^(k-1)) -- emitted radiance of the first
    # hit, then one term per further vertex (progressive_path.cpp:149, 175-184)
    for k in (1, 2, 4):
        r = osc.radiance(_params(pkg, max_depth=k, rr_depth=100), pix[:50000], np.arange(50000, dtype=np.uint32)).astype(np.float64)
        want_k = L * (1 - rho ** k) / (1 - rho)
        m, s = r.mean(0), r.std(0) / np.sqrt(50000)
        assert np.all(np.abs(m - want_k) <= 4 * s + 1e-3 * want_k), (k, m, want_k)


def test_guided_furnace(pkg, oracle):
    """One-sample MIS between the BSDF and a TRAINED guiding field must not bias the estimator: the furnace value L / (1 - rho)
    again, with directions drawn from the mixtures half of the time (field trained by two updates on this scene)."""
    sb, want = furnace_scene(pkg)
    osc = oracle.scene(sb)
    p = _params(pkg, max_depth=-1, rr_depth=5, guiding=1, guide_max_components=8, guide_max_cell_samples=2000)
    fld = oracle.field(8, (0, 0, 0), (1, 1, 1))
    sink = oracle.samples()
    for k in range(2):
        sink.clear()
        osc.render(p, 8 * k, 8, field=fld if k else None, sink=sink)
        fld.train_sink(sink, 4, 2000.0)
    assert fld.info()["cells"] > 1
    rng = np.random.RandomState(1)
    n = 200000
    pix = rng.randint(0, 256, n).astype(np.uint32)
    rad = osc.radiance(p, pix, 1000 + np.arange(n, dtype=np.uint32), field=fld).astype(np.float64)
    mean, sem = rad.mean(0), rad.std(0) / np.sqrt(n)
    assert np.all(np.abs(mean - want) <= 4 * sem + 1e-3 * want), (mean, want, sem)


@pytest.mark.parametrize("phase,g,method", [("isotropic", 0.0, "woodcock"), ("hg", 0.7, "woodcock"), ("hg", -0.3, "simpson")])
def test_volumetric_furnace(pkg, oracle, phase, g, method):
    """ProgressiveVolumetricPathTracer::Li (progressive_volpath.cpp:98-460): a heterogeneous medium that only scatters (albedo 1)
    inside the furnace box leaves the radiance at L / (1 - rho): free-flight sampling, phase-function sampling, emitter
    connections attenuated by the stochastic transmittance, the index-matched boundary and the surface part of the
    estimator all have to be consistent for that. The density varies in space (a smooth bump), both tracking methods.

    Finding (oracle/oracle_volpath.h, "REFERENCE QUIRK"): the reference is NOT unbiased here when next-event estimation is
    on. rayIntersectAndLookForEmitter hands pdfEmitterDirect the length of the LAST ray segment (after the index-matched
    boundary) instead of the distance from the vertex, so the MIS weights of the two strategies do not sum to one. Three legs:
    without NEE the reference algorithm returns the exact value; with NEE and the total distance (oracle test hook) it does
    too; with NEE as the reference does it, the result is several per cent too bright -- which this test pins, because the
    product reproduces the reference (parity), not the corrected estimator."""
    sb, want = furnace_scene(pkg, medium=(phase, g, method))
    osc = oracle.scene(sb)
    rng = np.random.RandomState(2)
    n = 120000
    pix = rng.randint(0, 256, n).astype(np.uint32)
    smp = np.arange(n, dtype=np.uint32)

    def mean_sem(**kw):
        r = osc.radiance(_params(pkg, max_depth=-1, rr_depth=5, volumetric=1, **kw), pix, smp).astype(np.float64)
        return r.mean(0), r.std(0) / np.sqrt(n)

    mean, sem = mean_sem(use_nee=0)                                  # the reference algorithm, phase / BSDF sampling only
    assert np.all(np.abs(mean - want) <= 4 * sem + 2e-3 * want), (mean, want, sem)
    before = oracle.lib.orc_debug_lookup_total_distance(1)           # NEE with the vertex-to-emitter distance in the MIS weight
    try:
        mean, sem = mean_sem()
    finally:
        oracle.lib.orc_debug_lookup_total_distance(before)
    assert np.all(np.abs(mean - want) <= 4 * sem + 2e-3 * want), (mean, want, sem)
    mean, sem = mean_sem()                                           # NEE as the reference does it: biased bright
    assert np.all(mean > 1.05 * want) and np.all(mean < 1.2 * want), (mean, want, sem)


@pytest.mark.parametrize("volumetric", [0, 1])
def test_furnace_with_a_glass_cube(pkg, oracle, volumetric):
    """A smooth dielectric neither absorbs nor emits: seen from outside, the furnace keeps L / (1 - rho) with a glass cube in
    it. Exercises SmoothDielectric::sample (dielectric.cpp:300-394): the Fresnel split between the two delta lobes, the
    radiance scaling factor^2 of transmitted radiance (:311-314) that must cancel over an entry / exit pair, bRec.eta in the
    Russian-roulette term eta^2 (progressive_path.cpp:296-306), total internal reflection chains inside the cube, and the
    rule that delta lobes take no next-event estimation and pass emitted radiance with weight one (:276-284). Both
    integrators (the volumetric one walks the same surfaces without any medium)."""
    sb, want = furnace_scene(pkg, glass=True)
    osc = oracle.scene(sb)
    rng = np.random.RandomState(4)
    n = 200000
    pix = rng.randint(0, 256, n).astype(np.uint32)
    rad = osc.radiance(_params(pkg, max_depth=-1, rr_depth=5, volumetric=volumetric), pix, np.arange(n, dtype=np.uint32)).astype(np.float64)
    mean, sem = rad.mean(0), rad.std(0) / np.sqrt(n)
    assert np.all(np.abs(mean - want) <= 4 * sem + 2e-3 * want), (mean, want, sem)


def test_two_lights_share_the_emitter_pmf(pkg, oracle):
    """Scene::sampleEmitterDirect picks an emitter from the discrete distribution over samplingWeight (scene.cpp:871-895,
    pmf.h:124-188, reusing the sample) and divides by its probability: two lights of different size and radiance add up to
    rho * (L1 F1 + L2 F2)."""
    S = pkg.scenes
    rho = 0.5
    sb = S.SceneBuilder(9, 9, spp=1)
    X = (1, 0, 0)
    sb.rectangle([S.scale(50, 50, 1), S.rotate(X, -90.0)], bsdf=sb.diffuse((rho, rho, rho)))
    lights = [((0.4, 0.3), (0.0, 1.2, 0.0), (4.0, 4.0, 4.0)), ((0.2, 0.6), (0.9, 2.0, -0.5), (1.0, 9.0, 2.0))]
    for (hx, hz), c, L in lights:
        sb.rectangle([S.scale(hx, hz, 1), S.rotate(X, 90.0), S.translate(*c)], bsdf=-1, radiance=L)
    target = np.array([0.2, 0.0, 0.1])
    sb.set_camera((3.0, 1.0, 2.5), tuple(target), (0, 1, 0), 0.05)
    osc = oracle.scene(sb)

    def signed_ff(x0, x1, z0, z1, h):  # rectangle [x0, x1] x [z0, z1] relative to the foot point, any position (superposition)
        f = lambda a, b: np.sign(a) * np.sign(b) * _corner_form_factor(abs(a), abs(b), h) if a != 0 and b != 0 else 0.0
        return f(x1, z1) - f(x0, z1) - f(x1, z0) + f(x0, z0)

    want = np.zeros(3)
    for (hx, hz), c, L in lights:
        want += rho * np.array(L) * signed_ff(c[0] - hx - target[0], c[0] + hx - target[0], c[2] - hz - target[2], c[2] + hz - target[2], c[1])
    n = 400000
    rad = osc.radiance(_params(pkg, max_depth=2), np.full(n, 40, np.uint32), np.arange(n, dtype=np.uint32)).astype(np.float64)
    mean, sem = rad.mean(0), rad.std(0) / np.sqrt(n)
    assert np.all(np.abs(mean - want) <= 4 * sem + 2e-3 * want), (mean, want, sem)


@pytest.mark.parametrize("volumetric", [0, 1])
def test_furnace_hide_emitters_and_strict_normals(pkg, oracle, volumetric):
    """hideEmitters drops only the emission of directly visible emitters ((!m_hideEmitters || scattered),
    progressive_path.cpp:167-169): the furnace loses exactly its first term, L / (1 - rho) - L. strictNormals changes nothing
    on flat-shaded geometry (geometric and shading normals agree, :175-184)."""
    sb, want = furnace_scene(pkg)
    osc = oracle.scene(sb)
    rng = np.random.RandomState(6)
    n = 150000
    pix = rng.randint(0, 256, n).astype(np.uint32)
    smp = np.arange(n, dtype=np.uint32)
    r = osc.radiance(_params(pkg, max_depth=-1, rr_depth=5, hide_emitters=1, volumetric=volumetric), pix, smp).astype(np.float64)
    mean, sem = r.mean(0), r.std(0) / np.sqrt(n)
    assert np.all(np.abs(mean - (want - 1.0)) <= 4 * sem + 1e-3), (mean, want - 1.0, sem)
    a = osc.radiance(_params(pkg, max_depth=-1, rr_depth=5, volumetric=volumetric), pix[:20000], smp[:20000])
    b = osc.radiance(_params(pkg, max_depth=-1, rr_depth=5, strict_normals=1, volumetric=volumetric), pix[:20000], smp[:20000])
    assert np.array_equal(a, b)


def test_furnace_image_is_flat_including_the_border(pkg, oracle):
    """The whole image pipeline on the furnace: tiles with filter borders merged into the film (imageproc.cpp:27-78,
    imageblock.h:131-197), Gaussian reconstruction filter, develop = colour / weight (fmtconv.cpp:978-1005). A constant
    radiance field must come out constant in every pixel -- also in the first and last rows / columns, where part of the
    filter footprint falls outside the image and only the weight division keeps the level."""
    from oracle_lib import develop

    sb, want = furnace_scene(pkg, res=40)                       # 40 x 40: two tiles in each direction (32-pixel tiles)
    osc = oracle.scene(sb)
    film, st = osc.render(_params(pkg, max_depth=-1, rr_depth=5), 0, 48)
    img = develop(film)
    assert img.shape == (40, 40, 3) and np.isfinite(img).all()
    assert abs(img.mean() - want) < 0.01
    border = np.concatenate([img[0].ravel(), img[-1].ravel(), img[:, 0].ravel(), img[:, -1].ravel()])
    inner = img[4:-4, 4:-4].ravel()
    assert abs(border.mean() - want) < 0.03 and abs(inner.mean() - want) < 0.01
    assert img.std() < 0.25                                     # per-pixel noise of 48 samples with std ~ 1 each, filtered
    # the weight channel: interior pixels collect the full normalised footprint of 48 samples per pixel
    assert abs(film[8:-8, 8:-8, 4].mean() / 48 - 1.0) < 0.02


def test_the_field_learns_where_the_light_is(pkg, oracle):
    """End-to-end sanity of the training loop (record path vertices -> weights = radiance / pdf -> binning -> weighted EM ->
    splits) on the form-factor scene rendered WITHOUT next-event estimation, where only sampled directions find the light:
    after a few updates the mixture of the cell under the light puts most of its mass into the light's solid angle
    (0.56 sr of the hemisphere's 6.28), and the guided estimator's variance drops well below the unguided one at equal samples
    while its mean stays at the closed-form value."""
    sb, centre, want = form_factor_scene(pkg)
    osc = oracle.scene(sb)
    K = 8
    p = _params(pkg, max_depth=2, use_nee=0, guiding=1, guide_max_components=K, guide_max_cell_samples=20000)
    fld = oracle.field(K, (0, 0, 0), (1, 1, 1))
    sink = oracle.samples()
    n = 60000
    pix = np.full(n, centre, np.uint32)
    for k in range(4):
        sink.clear()
        osc.radiance(p, pix, (k * n + np.arange(n)).astype(np.uint32), field=fld if k else None, sink=sink)
        fld.train_sink(sink, 4, 20000.0)
    # mass of the trained mixture inside the light's solid angle, seen from the floor point (0.5, 0, 0.1)
    m = 400
    gx = 0.3 + ((np.arange(m) + 0.5) / m - 0.5) * 1.6
    gz = -0.2 + ((np.arange(m) + 0.5) / m - 0.5) * 1.0
    X, Z = np.meshgrid(gx, gz, indexing="ij")
    v = np.stack([X.ravel() - 0.5, np.full(m * m, 1.5), Z.ravel() - 0.1], 1)
    r2 = (v * v).sum(1)
    w = (v / np.sqrt(r2)[:, None]).astype(np.float32)
    q = fld.pdf_sample(np.tile(np.float32([[0.5, 0.0, 0.1]]), (m * m, 1)), w, np.zeros((m * m, 3), np.float32))
    d_omega = (1.5 / np.sqrt(r2)) / r2 * (1.6 * 1.0 / (m * m))            # cos / r^2 dA
    mass = float((q["pdf"].astype(np.float64) * d_omega).sum())
    solid_angle = float(d_omega.sum())
    assert 0.5 < solid_angle < 0.62
    assert mass > 0.6, mass                                                # a uniform sphere would give solid_angle / 4 pi = 0.045
    # equal-sample comparison at the same pixel: guided (alpha = 0.5) against BSDF sampling only
    n2 = 200000
    smp = (10 * n + np.arange(n2)).astype(np.uint32)
    pix2 = np.full(n2, centre, np.uint32)
    g = osc.radiance(p, pix2, smp, field=fld).astype(np.float64)
    u = osc.radiance(_params(pkg, max_depth=2, use_nee=0), pix2, smp).astype(np.float64)
    for r in (g, u):
        mean, sem = r.mean(0), r.std(0) / np.sqrt(n2)
        assert np.all(np.abs(mean - want) <= 4 * sem + 2e-3 * want), (mean, want, sem)
    assert g.var(0).sum() < 0.5 * u.var(0).sum(), (g.var(0), u.var(0))


@pytest.mark.parametrize("guided_distance", [0, 1])
def test_guided_volumetric_furnace(pkg, oracle, guided_distance):
    """The guided volumetric path (direction guiding at medium and surface vertices; optionally guided free-flight sampling with
    weighted delta tracking) against the furnace value, with a trained field. Next-event estimation is off so that the
    reference's look-up quirk (test_volumetric_furnace) does not enter: what is tested is that the guided decisions --
    mixture / phase one-sample MIS, guided collision probabilities and their weights -- leave the estimator unbiased."""
    sb, want = furnace_scene(pkg, medium=("hg", 0.5, "woodcock"))
    osc = oracle.scene(sb)
    K = 8
    p = _params(pkg, max_depth=-1, rr_depth=5, volumetric=1, use_nee=0, guiding=1, guide_max_components=K,
                guide_max_cell_samples=3000, guided_distance=guided_distance)
    fld = oracle.field(K, (0, 0, 0), (1, 1, 1))
    sink = oracle.samples()
    for k in range(2):
        sink.clear()
        osc.render(p, 4 * k, 4, field=fld if k else None, sink=sink)
        fld.train_sink(sink, 4, 3000.0)
    assert fld.info()["cells"] > 1
    rng = np.random.RandomState(8)
    n = 150000
    pix = rng.randint(0, 256, n).astype(np.uint32)
    rad = osc.radiance(p, pix, 5000 + np.arange(n, dtype=np.uint32), field=fld).astype(np.float64)
    mean, sem = rad.mean(0), rad.std(0) / np.sqrt(n)
    assert np.all(np.abs(mean - want) <= 4 * sem + 2e-3 * want), (mean, want, sem)
