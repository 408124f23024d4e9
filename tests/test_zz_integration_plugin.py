"""The drop-in boundary from the reference's side: integration/b200guidedpath.cpp, a Mitsuba integrator plugin written against
the reference's own plugin interface (MTS_EXPORT_PLUGIN, Integrator::preprocess / render / cancel / postprocess) that forwards
to libb200pg.so's C-ABI. It is compiled against the reference build of oracle/_ref (integration/Makefile) and exercised here
through the reference itself: PluginManager::createObject loads plugins/b200guidedpath.so, Scene::render calls it, and the film
is read back from the reference's HDRFilm storage.

CPU: the plugin loads, parses the reference's parameter names, reads the scene XML -- and fails loudly at the device
(no CPU fallback). GPU: the film it puts into the reference's HDRFilm equals the film of a direct C-ABI render and, with
guiding off, the film the reference's own progressivepath renders from the same sample indices."""
import os

import numpy as np
import pytest

import ref_lib
from conftest import ROOT

PLUGIN = os.path.join(ROOT, "oracle", "_ref", "plugins", "b200guidedpath.so")
pytestmark = pytest.mark.skipif(not (ref_lib.available() and os.path.exists(PLUGIN)),
                                reason="oracle/_ref or the plugin not built (needs /root/reference at build time)")


def _scene(pkg, tmp_path, spp=4):
    sb = pkg.scenes.cornell_box(64, 64, spp=spp)
    xml = pkg.scenes.save_scene(sb, str(tmp_path))
    return sb, xml


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_plugin_loads_in_the_reference_and_fails_loudly_without_a_device(pkg, tmp_path):
    sb, xml = _scene(pkg, tmp_path)
    rs = ref_lib.RefScene(sb)
    p = pkg._abi.default_params()
    p.max_depth, p.guiding = 8, 0
    if _have_gpu():
        film, _ = rs.render_plugin(p, "b200guidedpath", xml)
        assert film[..., 4].sum() > 0
        return
    with pytest.raises(RuntimeError) as e:
        rs.render_plugin(p, "b200guidedpath", xml)
    assert "no CUDA device" in str(e.value) and "b200guidedpath" in str(e.value)  # the library's own message, through Log(EError)
    with pytest.raises(RuntimeError) as e:
        rs.render_plugin(p, "b200guidedpath", "")
    assert "no source file" in str(e.value)
    with pytest.raises(RuntimeError) as e:
        rs.render_plugin(p, "b200guidedpath", os.path.join(str(tmp_path), "missing.xml"))
    assert "b200guidedpath" in str(e.value)
    # the reference's own integrator still renders the same scene object afterwards
    film, _ = rs.render(p, 0, 1)
    assert film[..., 4].sum() > 0


def test_plugin_film_handoff_into_the_reference_film(pkg):
    """The one step of the plugin's render() that neither side of the device failure above reaches on a CPU-only machine: the
    GPU film (H x W x 5) becomes an ImageBlock with the film's filter border and goes through Film::put (hdrfilm.cpp:391-393);
    the reference's film storage must then hold exactly those numbers."""
    sb = pkg.scenes.cornell_box(48, 32, spp=1)
    rs = ref_lib.RefScene(sb)
    rgbaw = np.random.RandomState(1).rand(32, 48, 5).astype(np.float32)
    assert np.array_equal(rs.plugin_put_film(rgbaw), rgbaw)


def _gpu_body(tmp_path):
    """Runs in a child process (see the test below): Scene::render of the reference with the GPU integrator plugged in."""
    from conftest import load_package

    pkg = load_package()
    from b200pg import api

    sb, xml = _scene(pkg, tmp_path)
    rs = ref_lib.RefScene(sb)
    p = pkg._abi.default_params()
    p.max_depth, p.guiding = 8, 0
    film, sec = rs.render_plugin(p, "b200guidedpath", xml)        # Scene::render -> plugin -> b200pg_render -> Film::put
    sc = api.Scene.load_xml(xml)
    it = api.Integrator(sc, p)
    it.render()
    direct = it.film()
    it.close()
    np.testing.assert_allclose(film[..., 4], direct[..., 4], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(film[..., :3], direct[..., :3], rtol=1e-4, atol=1e-3)
    own, _ = rs.render(p, 0, 4)                                    # the reference's progressivepath, same sample indices
    np.testing.assert_allclose(film[..., 4], own[..., 4], rtol=1e-4, atol=1e-4)
    dev_g = film[..., :3] / np.maximum(film[..., 4:5], 1e-20)
    dev_r = own[..., :3] / np.maximum(own[..., 4:5], 1e-20)
    assert np.abs(dev_g - dev_r).mean() / dev_r.mean() < 5e-3
    pg = pkg._abi.default_params()
    pg.max_depth, pg.guiding, pg.training_progressions = 8, 1, 2
    gfilm, _ = rs.render_plugin(pg, "b200guidedpath", xml)        # the guided integrator through the same door
    dev = gfilm[..., :3] / np.maximum(gfilm[..., 4:5], 1e-20)
    assert np.isfinite(dev).all() and abs(dev.mean() - dev_r.mean()) < 0.2 * dev_r.mean()  # 2 - 4 spp at 64 x 64: noise
    print("plugin ok: film through the reference's render loop == direct render; mean %.4f vs the reference's own %.4f; guided %.4f"
          % (dev_g.mean(), dev_r.mean(), dev.mean()))


@pytest.mark.gpu
def test_reference_render_loop_with_the_gpu_integrator(tmp_path):
    """In a child process: the reference's thread / scheduler singletons and the CUDA runtime share it, and whatever happens
    there must show up as one failed test, not as the end of the test run."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.abspath(__file__), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "plugin ok" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])


if __name__ == "__main__":
    import pathlib
    import sys

    _gpu_body(pathlib.Path(sys.argv[1]))
