"""GPU parity of branches the main suites do not reach ("untested branches"):
  * a finite maxComponentValue (the clamp of ProgressiveMonteCarloIntegrator::renderBlock, progressiveintegrator.cpp:274-277);
  * a triangle-mesh AREA LIGHT: Shape::samplePosition of a TriMesh (trimesh.cpp:412-423, triangle.cpp:24-59) in next-event
    estimation and its pdf in the emitter-hit MIS weight;
  * obj / ply meshes loaded by the XML reader, shaded with the smooth normals TriMesh::computeNormals gives them
    (trimesh.cpp:631-668), including the geometric-normal flip of skdtree.h:343-428 -- rendered on the GPU and by the oracle from the
    SAME flat description (what the reader produced);
  * b200pg_render with a device list (in-library multi-device render) and the cancel flag no longer sticking.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(pkg):
    from b200pg import api as _api

    return _api


def test_finite_max_component_value(pkg, api, oracle):
    from oracle_lib import develop

    sb = pkg.scenes.cornell_caustic(96, 96, spp=8)
    osc = oracle.scene(sb)
    imgs = {}
    for clamp in (float("inf"), 0.75):
        p = api.default_params()
        p.max_depth = 8
        p.max_component_value = clamp
        it = api.Integrator(api.Scene.from_builder(sb), p)
        it.progression(0, 16)
        img_g = it.develop()
        film_o, _ = osc.render(p, 0, 16)
        img_o = develop(film_o)
        assert np.abs(img_g - img_o).sum() / np.abs(img_o).sum() < 3e-3, clamp
        imgs[clamp] = img_g
        # per-sample radiance is NOT clamped (the clamp belongs to renderBlock, not to Li)
        if clamp < 1:
            pix = np.arange(0, 96 * 96, 7, dtype=np.uint32)
            L = it.k_radiance(pix, np.zeros_like(pix))
            assert L.max() > clamp
    # the clamp removes energy (fireflies of the caustic paths) and bounds every developed pixel
    assert imgs[0.75].mean() < 0.9 * imgs[float("inf")].mean()
    assert imgs[0.75].max() <= 0.75 * (1 + 1e-4)


def _write_meshes(tmp_path, S):
    # a bumpy patch (obj, no normals -> smooth normals) and a quad light made of two triangles (ply)
    P, _, T = S.heightfield_mesh(n=24, seed=3, amp=0.6)
    P = P.astype(np.float32) * np.float32(0.9)
    with open(tmp_path / "patch.obj", "w") as f:
        for p in P:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        for t in T:
            f.write("f %d %d %d\n" % tuple(t + 1))
    Q = np.array([[-0.3, 1.2, -0.3], [0.3, 1.2, -0.3], [0.3, 1.2, 0.3], [-0.3, 1.3, 0.3]], np.float32)  # not planar: two normals
    F = np.array([[0, 1, 2], [0, 2, 3]], np.int32)  # facing down (-y)
    hdr = "ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n" \
          "element face 2\nproperty list uchar int vertex_indices\nend_header\n"
    with open(tmp_path / "light.ply", "w") as f:
        f.write(hdr)
        for q in Q:
            f.write("%.9g %.9g %.9g\n" % tuple(q))
        for t in F:
            f.write("3 %d %d %d\n" % tuple(t))


_SCENE = '''<?xml version="1.0"?>
<scene version="0.6.0">
  <integrator type="progressivepath"><integer name="maxDepth" value="5"/></integrator>
  <sensor type="perspective"><float name="fov" value="50"/>
    <transform name="toWorld"><lookat origin="0, 1.6, 2.6" target="0, 0.1, 0" up="0, 1, 0"/></transform>
    <sampler type="independent"><integer name="sampleCount" value="8"/></sampler>
    <film type="hdrfilm"><integer name="width" value="96"/><integer name="height" value="96"/><boolean name="banner" value="false"/></film>
  </sensor>
  <shape type="obj"><string name="filename" value="patch.obj"/>
    <bsdf type="roughplastic"><rgb name="diffuseReflectance" value="0.5, 0.3, 0.2"/><float name="alpha" value="0.2"/></bsdf></shape>
  <shape type="ply"><string name="filename" value="light.ply"/>%s
    <emitter type="area"><rgb name="radiance" value="14, 12, 9"/></emitter></shape>
  <shape type="rectangle"><transform name="toWorld"><rotate x="1" angle="-90"/><scale value="3"/><translate y="-0.4"/></transform>
    <bsdf type="diffuse"><rgb name="reflectance" value="0.6"/></bsdf></shape>
</scene>
'''


@pytest.mark.parametrize("face_normals", [False, True])
def test_mesh_light_and_loaded_meshes_with_smooth_normals(pkg, api, oracle, tmp_path, face_normals):
    from oracle_lib import OracleScene, develop

    _write_meshes(tmp_path, pkg.scenes)
    xml = tmp_path / "scene.xml"
    xml.write_text(_SCENE % ('<boolean name="faceNormals" value="true"/>' if face_normals else ""))
    sc = api.Scene.load_xml(str(xml))
    d = sc.desc
    assert d.n_shapes == 3 and d.n_emitters == 1
    assert bool(d.shapes[0].normals)                       # obj without vn: smooth normals computed by the loader
    assert bool(d.shapes[1].normals) != face_normals       # the light: smooth unless faceNormals
    p = sc.integrator_params()
    assert p.max_depth == 5
    it = api.Integrator(sc, p)
    osc = OracleScene.from_desc(oracle, d, keep=sc)
    rng = np.random.RandomState(9)
    n = 60000
    pix = rng.randint(0, 96 * 96, n).astype(np.uint32)
    smp = rng.randint(0, 64, n).astype(np.uint32)
    Lo, Lg = osc.radiance(p, pix, smp), it.k_radiance(pix, smp)
    assert (Lo.max(1) > 0).mean() > 0.25                   # the mesh light is found (NEE + hits)
    err = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-3)
    assert (err > 1e-3).mean() < 5e-3, (err > 1e-3).mean()
    assert abs(Lg.mean() - Lo.mean()) <= 3e-3 * Lo.mean()
    # NEE off: the light is only found by BSDF sampling -> emitter-hit radiance of a mesh light, same estimator mean
    p2 = sc.integrator_params()
    p2.use_nee = 0
    it2 = api.Integrator(sc, p2)
    Lo2, Lg2 = osc.radiance(p2, pix, smp), it2.k_radiance(pix, smp)
    err2 = np.abs(Lo2 - Lg2).max(1) / (np.abs(Lo2).max(1) + 1e-3)
    assert (err2 > 1e-3).mean() < 5e-3
    # whole image through the progression path as well
    it.progression(0, 8)
    film_o, st = osc.render(p, 0, 8)
    img_g, img_o = it.develop(), develop(film_o)
    assert np.abs(img_g - img_o).sum() / np.abs(img_o).sum() < 5e-3


def test_cancel_does_not_stick_and_render_takes_a_device_list(pkg, api):
    sb = pkg.scenes.cornell_box(64, 64, spp=8)
    p = api.default_params()
    p.max_depth = 6
    p.samples_per_progression = 2
    it = api.Integrator(api.Scene.from_builder(sb), p)
    it.cancel()
    it.render(devices=[0])          # a cancel request before the call must not empty it
    st = it.stats()
    assert st["paths"] == 64 * 64 * 8 and st["progressions_done"] == 4
    with pytest.raises(api.B200pgError, match="device"):
        it.render(devices=[7])       # the integrator lives on device 0
    with pytest.raises(api.B200pgError, match="twice|device"):
        it.render(devices=[0, 0])


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_in_library_multi_device_render_matches_single_device(pkg, api):
    """b200pg_render(integ, 2, {0, 1}): same samples as the one-device render (sample blocks are keyed by their index), fields
    trained from the union of both devices' samples, films merged into the calling integrator."""
    sb = pkg.scenes.cornell_caustic(128, 128, spp=16)

    def make():
        p = api.default_params()
        p.max_depth, p.samples_per_progression, p.guiding, p.guide_max_components = 8, 2, 1, 8
        p.training_progressions, p.guide_train_discard_film = 4, 1
        return api.Integrator(api.Scene.from_builder(sb), p)
    a, b = make(), make()
    a.render()
    b.render(devices=[0, 1])
    sa, sb_ = a.stats(), b.stats()
    assert sa["paths"] == sb_["paths"] == 128 * 128 * 16
    ia, ib = a.develop(), b.develop()
    # the training schedules differ (one update per TWO sample blocks on two devices), so the guided estimators differ: both
    # are unbiased -> equal means within noise, and both films hold the same number of samples
    assert abs(ia.mean() - ib.mean()) <= 0.03 * ia.mean()
    fa, fb = a.film(), b.film()
    assert abs(fa[..., 4].sum() - fb[..., 4].sum()) <= 1e-4 * fa[..., 4].sum()
    assert b.stats()["guide_cells"] >= 1
    # a second call on the same integrator works (peers are disconnected again)
    b.render(devices=[0, 1])
    assert b.stats()["paths"] == 2 * 128 * 128 * 16
