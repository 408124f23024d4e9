"""Host-side checks that run without a GPU: the C-ABI library loads and exports every symbol include/b200pg.h
declares, the host scene compiler accepts/rejects descriptions like the reference's constructors do, the product's
rough-transmittance reduction agrees with the oracle's, and compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def api(pkg):
    from b200pg import api as _api

    if not os.path.exists(_api.LIB_PATH):
        import subprocess

        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "mitsuba-path-guiding_b200", "csrc")])
    return _api


def test_every_declared_symbol_is_exported(api):
    hdr = open(os.path.join(ROOT, "include", "b200pg.h")).read()
    names = sorted(set(re.findall(r"\b(b200pg_\w+)\s*\(", hdr)))
    assert len(names) >= 30
    lib = C.CDLL(api.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.b200pg_version() == 100


def test_struct_layouts_match_header(api, pkg):
    """Round-trip through the library: what Python writes into the ctypes mirror is what C++ reads."""
    sb = pkg.scenes.cornell_caustic(96, 64, spp=7, seed=99)
    sc = api.Scene.from_builder(sb)
    d = sc.desc
    assert (d.film.width, d.film.height, d.sample_count, d.seed) == (96, 64, 7, 99)
    assert d.n_shapes == len(sb.shapes) and d.n_emitters == 1
    # Shape::configure defaults were applied (shape.cpp:48-70): the emitter-only rectangle got a black diffuse BSDF
    light = [i for i, s in enumerate(sb.shapes) if s["emitter"] >= 0][0]
    b = d.bsdfs[d.shapes[light].bsdf]
    assert b.type == pkg._abi.BSDF_DIFFUSE and list(b.reflectance) == [0, 0, 0]
    glass = [i for i in range(d.n_bsdfs) if d.bsdfs[i].type == pkg._abi.BSDF_DIELECTRIC]
    assert glass and abs(d.bsdfs[glass[0]].int_ior - 1.5) < 1e-7
    p = api.default_params()
    assert (p.max_depth, p.rr_depth, p.samples_per_progression, p.use_nee) == (-1, 5, 1, 1)  # integrator.cpp:195-230
    assert p.max_component_value == float("inf")


def test_scene_errors(api, pkg):
    S = pkg.scenes
    sb = S.SceneBuilder(8, 8)
    sb.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0)
    with pytest.raises(api.B200pgError, match="no shapes"):
        api.Scene.from_builder(sb)
    sb.rectangle([np.array([[1, 0.5, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)])
    with pytest.raises(api.B200pgError, match="shear"):  # rectangle.cpp:105-106
        api.Scene.from_builder(sb)
    sb = S.SceneBuilder(8, 8)
    sb.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0)
    sb.rectangle([S.scale(1, 1, 1)], bsdf=sb.roughplastic(alpha=0.1, int_ior=1.0, ext_ior=1.0))
    with pytest.raises(api.B200pgError, match="refraction"):  # roughplastic.cpp:209-211
        api.Scene.from_builder(sb)
    sb = S.SceneBuilder(8, 8)
    sb.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0)
    sb.rectangle([S.scale(1, 1, 1)], bsdf=sb.roughplastic(alpha=5.0))
    with pytest.raises(api.B200pgError, match="roughness"):  # rtrans.h checkAlpha
        api.Scene.from_builder(sb)
    sb = S.SceneBuilder(8, 8)
    sb.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0)
    sb.trimesh(np.zeros((3, 3), np.float32), np.array([[0, 1, 5]], np.uint32))
    with pytest.raises(api.B200pgError, match="index out of range"):
        api.Scene.from_builder(sb)


def test_rtrans_reduction_matches_oracle(api, pkg, oracle):
    """Product loader (host_scene.cpp) vs oracle restatement of rtrans.h:292-388 on the same packed table."""
    from b200pg import rtrans

    S = pkg.scenes
    for distr, eta_pair, alpha in (("beckmann", (1.49, 1.000277), 0.7), ("ggx", (1.49, 1.000277), 0.2),
                                   ("beckmann", (1.9, 1.0), 0.05), ("ggx", (1.33, 1.0), 1.5)):
        sb = S.SceneBuilder(8, 8)
        sb.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0)
        sb.rectangle([S.scale(1, 1, 1)], bsdf=sb.roughplastic(alpha=alpha, int_ior=eta_pair[0], ext_ior=eta_pair[1], distribution=distr))
        sc = api.Scene.from_builder(sb)
        b = sc.desc.bsdfs[0]
        eta = float(np.float32(eta_pair[0]) / np.float32(eta_pair[1]))
        ext, ed, idf = oracle.rtrans_reduce(rtrans.load_packed(distr), eta, alpha)
        np.testing.assert_allclose(np.array(b.rt_ext_trans[:]), ext, rtol=2e-6, atol=2e-7)
        assert abs(b.rt_ext_diff - ed) < 1e-6 and abs(b.rt_int_diff - idf) < 1e-6
        assert 0.0 <= ext.min() and ext.max() <= 1.0 + 1e-6
        # physical anchor: at normal incidence the smooth limit is 1 - F(eta); rough surfaces transmit slightly less
        F = ((eta - 1) / (eta + 1)) ** 2
        assert ext[-1] <= 1 - F + 5e-3 and ext[-1] > 0.5


def test_no_cpu_fallback(api, pkg):
    """Without a CUDA device integrator creation must fail loudly (the product path has no CPU route)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sc = api.Scene.from_builder(pkg.scenes.cornell_box(16, 16))
    with pytest.raises(api.B200pgError, match="CUDA"):
        api.Integrator(sc, api.default_params())


def test_header_is_plain_c_and_links_from_c(pkg, tmp_path):
    """The boundary is a C ABI: include/b200pg.h compiles as strict C99 (no C++ or torch types in any signature) and a C
    program links against libb200pg.so and calls it (here: the entry points that need no GPU)."""
    import shutil
    import subprocess

    from conftest import PKG_DIR, ROOT

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "b200pg.h"\n'
                   'int main(void) {\n'
                   '    B200pgIntegratorParams p;\n'
                   '    b200pg_integrator_params_default(&p);\n'
                   '    char err[256];\n'
                   '    void *s = b200pg_scene_load_xml("/nonexistent.xml", NULL, err, sizeof err);\n'
                   '    printf("%d %d %d %s\\n", b200pg_version(), p.rr_depth, s == NULL, err);\n'
                   '    return 0;\n}\n')
    exe = tmp_path / "t_c"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-L", PKG_DIR, "-lb200pg", "-Wl,-rpath," + PKG_DIR, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    ver, rr, null, msg = out.stdout.split(" ", 3)
    assert int(ver) >= 100 and int(rr) == 5 and null == "1" and "cannot open" in msg
