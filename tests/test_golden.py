"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py -- see its header for what they
are and are not).

CPU (`-m "not gpu"`): the oracle reproduces every fixture (integers exactly; floats to 1e-6 relative -- the same code on
another host may contract / vectorise libm calls differently; the tile-merge order of the threaded render to 1e-5).
GPU (`-m gpu`): the CUDA path, through the C-ABI, against the same fixtures with the tolerances of tests/test_gpu_parity.py
and tests/test_gpu_guiding.py (bit-exact indices / binning; 1e-5 relative for pdfs and EM statistics)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_golden as mg  # noqa: E402

FILES = ("cornell", "caustic", "bsdf", "medium", "guiding")


@pytest.fixture(scope="module")
def gold():
    return {f: dict(np.load(os.path.join(HERE, "golden", f + ".npz"))) for f in FILES}


def test_fixtures_are_committed_and_small(gold):
    total = sum(os.path.getsize(os.path.join(HERE, "golden", f + ".npz")) for f in FILES)
    assert total < 4 << 20
    assert gold["cornell"]["prim"].dtype == np.uint32 and gold["guiding"]["field"][0] == 0x47554944


def test_oracle_reproduces_the_fixtures(pkg, oracle, gold):
    out = mg.generate(pkg, oracle)
    assert sorted(out) == sorted(FILES)
    for name, d in out.items():
        assert sorted(d) == sorted(gold[name]), name
        for k, v in d.items():
            ref = gold[name][k]
            assert ref.shape == v.shape and ref.dtype == v.dtype, (name, k)
            if v.dtype.kind in "ub":
                if k in ("field", "field_after"):  # snapshots hold float bit patterns
                    np.testing.assert_allclose(v.view(np.float32)[8:], ref.view(np.float32)[8:], rtol=1e-6, atol=1e-7, err_msg=k)
                    assert np.array_equal(v[:8], ref[:8])
                elif k == "render_counters":
                    assert np.array_equal(v, ref)
                else:
                    assert np.array_equal(v, ref), (name, k)
            else:
                tol = 1e-5 if k == "render_film" else 1e-6
                np.testing.assert_allclose(v, ref, rtol=tol, atol=tol, err_msg="%s/%s" % (name, k))


# ---------------------------------------------------------------------------------------------------------------------
#  GPU
# ---------------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(pkg):
    from b200pg import api as _api

    return _api


def _check_hits(tuv_o, prim_o, tuv_g, prim_g, max_mismatch=0):
    mism = prim_o != prim_g
    assert mism.sum() <= max_mismatch
    m = (prim_o != 0xFFFFFFFF) & ~mism
    rel = np.abs(tuv_o[m, 0] - tuv_g[m, 0]) / (1 + np.abs(tuv_o[m, 0]))
    duv = np.abs(tuv_o[m, 1:] - tuv_g[m, 1:]).max(1)
    assert np.quantile(rel, 0.999) <= 2e-6 and np.quantile(duv, 0.999) <= 2e-5
    assert rel.max() <= 2e-3 and duv.max() <= 5e-3


@gpu
@pytest.mark.parametrize("name", ["cornell", "caustic"])
def test_gpu_trace_radiance_film_against_fixtures(name, pkg, api, gold):
    g = gold[name]
    sb = mg.cases(pkg)[name]()
    it = api.Integrator(api.Scene.from_builder(sb), mg.params(pkg))
    tuv, prim = it.k_trace(g["rays"])
    _check_hits(g["tuv"], g["prim"], tuv, prim)
    tuv2, prim2 = it.k_trace(g["rays2"])
    _check_hits(g["tuv2"], g["prim2"], tuv2, prim2, max_mismatch=2)
    _, occ = it.k_trace(g["rays3"], shadow=True)
    assert ((occ != 0xFFFFFFFF) != g["occluded"]).sum() <= 2
    got = it.k_radiance(g["pixel"], g["sample"])
    want = g["radiance"]
    err = np.abs(got - want).max(1) / (np.abs(want).max(1) + 1e-3)
    assert (err > 1e-3).mean() < 3e-3
    if name == "cornell":
        it.film_clear()
        it.k_film_splat(g["splat_pos"], g["splat_rgb"])
        film = it.film()
        np.testing.assert_allclose(film[..., 4], g["film"][..., 4], rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(film[..., :3], g["film"][..., :3], rtol=2e-5, atol=1e-4)
        it.film_clear()
        s0 = it.stats()
        it.progression(0, 2)
        s1 = it.stats()
        f = it.film()
        np.testing.assert_allclose(f[..., 4], g["render_film"][..., 4], rtol=1e-4, atol=1e-4)
        dev_g = f[..., :3] / np.maximum(f[..., 4:5], 1e-20)
        dev_o = g["render_film"][..., :3] / np.maximum(g["render_film"][..., 4:5], 1e-20)
        assert np.abs(dev_g - dev_o).mean() / dev_o.mean() < 5e-3
        for i, key in enumerate(("paths", "normal_rays", "shadow_rays", "path_length_sum")):
            a, b = s1[key] - s0[key], int(g["render_counters"][i])
            assert abs(a - b) <= 3e-3 * b + 2, key
    it.close()


@gpu
def test_gpu_bsdfs_against_fixtures(pkg, api, gold):
    from bsdf_cases import bsdf_scene

    sb, idx = bsdf_scene()
    it = api.Integrator(api.Scene.from_builder(sb), mg.params(pkg))
    g = gold["bsdf"]
    for name, i in idx.items():
        o = {k: g["%s/%s" % (name, k)] for k in ("wi", "wo_in", "u", "eval", "pdf", "wo", "weight", "spdf", "flags")}
        r = it.k_bsdf(i, o["wi"], o["wo_in"], o["u"])
        for key in ("eval", "pdf"):
            assert np.all(np.abs(o[key] - r[key]) <= 1e-5 * np.maximum(np.abs(o[key]), 1e-3)), (name, key)
        same = o["flags"] == r["flags"]
        assert same.mean() > 0.998, name
        ok = same & (o["spdf"] > 0)
        assert np.abs(o["wo"][ok] - r["wo"][ok]).max() <= 2e-3, name
    it.close()


@gpu
def test_gpu_medium_against_fixtures(pkg, api, gold):
    g = gold["medium"]
    sb = mg.cases(pkg)["medium"]()
    p = mg.params(pkg, volumetric=1)
    it = api.Integrator(api.Scene.from_builder(sb), p)
    np.testing.assert_allclose(it.k_grid_lookup(0, g["points"]), g["density"], rtol=1e-5, atol=1e-6)
    t, tr, wo, pdf = it.k_medium_sample(0, g["rays"])
    # free-flight sampling is a chain of Woodcock decisions on one RNG stream: t = inf marks a ray that left the medium
    # (or reached maxt) without an interaction; same bars as tests/test_gpu_medium.py
    to, tro = g["t"], g["transmittance"]
    same = np.isfinite(to) == np.isfinite(t)
    assert same.mean() > 0.998
    m = same & np.isfinite(to)
    assert m.mean() > 0.1
    rel = np.abs(to[m] - t[m]) / np.maximum(np.abs(to[m]), 1e-3)
    assert (rel > 1e-5).mean() < 2e-3
    assert (tro != tr).mean() < 2e-3
    ok = same & (tro == tr)
    np.testing.assert_allclose(wo[ok], g["wo"][ok], atol=2e-5)
    np.testing.assert_allclose(pdf[ok], g["pdf"][ok], rtol=2e-4)
    it.close()


@gpu
def test_gpu_guiding_against_fixtures(pkg, api, gold):
    g = gold["guiding"]
    K = int(g["K"][0])
    sb = mg.cases(pkg)["caustic"]()
    it = api.Integrator(api.Scene.from_builder(sb), mg.params(pkg, guiding=1, guide_max_components=K, guide_max_cell_samples=3000))
    it.field_load(g["field"])
    q = it.k_vmm_pdf_sample(g["qpos"], g["qdir"], g["qu"])
    assert np.array_equal(q["cell"], g["q_cell"])                                     # indexing: bit-exact
    assert np.all(np.abs(q["pdf"] - g["q_pdf"]) <= 1e-5 * np.maximum(g["q_pdf"], 1e-3))
    assert np.abs(q["dir"] - g["q_dir"]).max() <= 1e-5
    assert np.all(np.abs(q["spdf"] - g["q_spdf"]) <= 1e-5 * np.maximum(g["q_spdf"], 1e-2))
    nc = int(g["field"][2])
    cell, perm, off = it.k_bin_samples(g["qpos"], nc)
    assert np.array_equal(cell, g["bin_cell"]) and np.array_equal(perm, g["bin_perm"]) and np.array_equal(off, g["bin_offsets"])
    s = dict(pos=g["s_pos"], dir=g["s_dir"], weight=g["s_weight"], pdf=g["s_pdf"], dist=g["s_dist"])
    st = it.k_em_step(s, 0, nc, K)                                                     # E-step statistics only
    ref = g["estep_stats"]
    scale = np.maximum(np.abs(ref).max(1, keepdims=True), 1e-6)
    assert np.all(np.abs(st - ref) <= 1e-5 * scale + 1e-5 * np.abs(ref))
    it.k_em_step(s, 4, nc, K)                                                          # the whole update: refit + split
    after, want = it.field_snapshot(), g["field_after"]
    assert np.array_equal(after[:8], want[:8])                                         # same tree size
    nn, nc2 = int(want[1]), int(want[2])
    na, nb = after[8:8 + 4 * nn].reshape(nn, 4), want[8:8 + 4 * nn].reshape(nn, 4)
    assert np.array_equal(na[:, [0, 2, 3]], nb[:, [0, 2, 3]])                          # topology: bit-exact
    np.testing.assert_allclose(na[:, 1].view(np.float32), nb[:, 1].view(np.float32), rtol=2e-6, atol=1e-6)
    o = 8 + 4 * nn + 8 * nc2
    la, lb = after.view(np.float32)[o:].reshape(-1, 12), want.view(np.float32)[o:].reshape(-1, 12)
    assert np.abs(la[:, 0] - lb[:, 0]).max() <= 1e-5                                   # mixture weights
    heavy = lb[:, 0] > 1e-3
    assert np.abs(la[heavy, 1:4] - lb[heavy, 1:4]).max() <= 5e-5                       # mean directions
    it.close()
