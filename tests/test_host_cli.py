"""b200pg-render (the `mitsuba` CLI subset, src/mitsuba/mitsuba.cpp:52-91) on a box without a GPU: argument handling and the
multi-GPU launcher's process plumbing (-p n forks one worker per GPU; a worker that cannot start must take the whole job
down with a non-zero exit code and the library's error text -- there is no CPU fallback to hide behind)."""
import os
import subprocess

import pytest

from conftest import PKG_DIR

EXE = os.path.join(PKG_DIR, "b200pg-render")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def xml(pkg, tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    return pkg.scenes.save_scene(pkg.scenes.cornell_box(32, 32, spp=4), str(d))


def _run(*args):
    return subprocess.run([EXE, *args], capture_output=True, text=True, timeout=120)


def test_usage_and_argument_errors(xml):
    r = _run("-h")
    assert r.returncode == 0 and "-p count" in r.stdout and "-r sec" in r.stdout and "-D key=val" in r.stdout
    assert _run().returncode == 1                              # no scene
    r = _run("-p", "0", xml)
    assert r.returncode == 1 and "GPU count" in r.stderr
    r = _run("--bogus", xml)
    assert r.returncode == 1 and "unknown option" in r.stderr
    r = _run("/nonexistent/scene.xml")
    assert r.returncode == 2 and "Error" in r.stderr


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour WITHOUT a CUDA device")
@pytest.mark.parametrize("extra", [(), ("-r", "1")])
def test_multi_gpu_launcher_fails_loudly_without_devices(xml, tmp_path, extra):
    out = str(tmp_path / "o.pfm")
    r = _run("-p", "2", *extra, "-o", out, xml)
    assert r.returncode == 2
    assert r.stderr.count("no CUDA device available (this library has no CPU path)") == 2   # both workers said why
    assert "a worker failed during start-up" in r.stderr
    assert not os.path.exists(out)
    r1 = _run("-o", out, xml)
    assert r1.returncode == 2 and "no CPU path" in r1.stderr
