"""Denoiser feature buffers (src/librender/denoiser.cpp:138-144: per-pixel running means of sample colour, albedo and
normal). CPU: the oracle's statement against independent recomputations; GPU: the CUDA path (k_features / k_feature_color,
b200pg_features_read / _write) against the oracle on the same samples."""
import struct

import numpy as np
import pytest


def _params(pkg, **kw):
    p = pkg._abi.default_params()
    p.max_depth = 8
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_oracle_features_are_running_means_of_the_pixel_samples(pkg, oracle):
    sb = pkg.scenes.cornell_box(48, 48, spp=4)
    osc = oracle.scene(sb)
    p = _params(pkg)
    f = osc.features(p, 0, 3)
    assert f.shape == (48, 48, 10) and (f[..., 9] == 3).all()
    # colour = mean of the per-sample radiance of the same (pixel, sample) streams
    pix = np.repeat(np.arange(48 * 48, dtype=np.uint32), 3)
    smp = np.tile(np.arange(3, dtype=np.uint32), 48 * 48)
    rad = osc.radiance(p, pix, smp).reshape(48, 48, 3, 3).mean(2)
    np.testing.assert_allclose(f[..., :3], rad, rtol=2e-6, atol=1e-6)
    # a second call continues the running mean (Denoiser::add): 3 + 1 samples == 4 samples at once
    g = osc.features(p, 3, 1, out=f.copy())
    h = osc.features(p, 0, 4)
    np.testing.assert_allclose(g, h, rtol=1e-5, atol=1e-6)
    # albedo / normal: the left wall is red (.63, .065, .05) with normal +x, the floor white (.725, .71, .68) with +y
    left, floor = h[24, 2], h[46, 24]
    np.testing.assert_allclose(left[3:6], [0.63, 0.065, 0.05], atol=1e-6)
    np.testing.assert_allclose(left[6:9], [1, 0, 0], atol=1e-5)
    np.testing.assert_allclose(floor[3:6], [0.725, 0.71, 0.68], atol=1e-6)
    np.testing.assert_allclose(floor[6:9], [0, 1, 0], atol=1e-5)
    n = np.linalg.norm(h[..., 6:9], axis=2)
    assert (n <= 1 + 1e-5).all() and np.median(n) > 0.999


def _read_exr_channels(path):
    """Minimal reader for the uncompressed scanline float32 files b200pg_features_write produces."""
    with open(path, "rb") as f:
        data = f.read()
    assert struct.unpack_from("<I", data, 0)[0] == 20000630
    pos, chans, W, H = 8, [], None, None
    while data[pos] != 0:
        end = data.index(b"\0", pos)
        name = data[pos:end].decode()
        pos = end + 1
        end = data.index(b"\0", pos)
        pos = end + 1
        size = struct.unpack_from("<i", data, pos)[0]
        pos += 4
        if name == "channels":
            q = pos
            while data[q] != 0:
                e = data.index(b"\0", q)
                chans.append(data[q:e].decode())
                assert struct.unpack_from("<i", data, e + 1)[0] == 2  # FLOAT
                q = e + 1 + 16
        if name == "dataWindow":
            x0, y0, x1, y1 = struct.unpack_from("<4i", data, pos)
            W, H = x1 - x0 + 1, y1 - y0 + 1
        if name == "compression":
            assert data[pos] == 0
        pos += size
    pos += 1 + 8 * H
    out = np.zeros((H, W, len(chans)), np.float32)
    for y in range(H):
        yy, sz = struct.unpack_from("<2i", data, pos)
        pos += 8
        line = np.frombuffer(data, np.float32, W * len(chans), pos).reshape(len(chans), W)
        out[yy] = line.T
        pos += sz
    return chans, out


@pytest.mark.gpu
def test_gpu_features_match_the_oracle(pkg, oracle, tmp_path):
    from b200pg import api

    sb = pkg.scenes.cornell_caustic(64, 64, spp=4)
    p = _params(pkg)
    it = api.Integrator(api.Scene.from_builder(sb), p)
    with pytest.raises(api.B200pgError):
        it.features()                                   # not enabled yet: fails loudly
    it.set_option("feature_buffers", 1)
    it.progression(0, 3)
    it.progression(3, 1)                                # sums continue across progressions
    got = it.features()
    want = oracle.scene(sb).features(p, 0, 4)
    assert (got[..., 9] == 4).all()
    # first-hit albedo / normal: same hits on both sides except for samples within rounding of a silhouette
    bad = (np.abs(got[..., 3:9] - want[..., 3:9]).max(2) > 1e-5)
    assert bad.mean() < 2e-3
    # colour: the same per-sample radiance up to a few flipped paths (tests/test_gpu_parity.py bar)
    err = np.abs(got[..., :3] - want[..., :3]).max(2) / (np.abs(want[..., :3]).max(2) + 1e-3)
    assert (err > 1e-3).mean() < 5e-3
    assert abs(got[..., :3].mean() - want[..., :3].mean()) < 3e-3 * want[..., :3].mean()
    # the film is untouched by the option: same image as an integrator without feature buffers
    it2 = api.Integrator(api.Scene.from_builder(sb), p)
    it2.progression(0, 4)
    np.testing.assert_allclose(it.film(), it2.film(), rtol=1e-4, atol=1e-4)
    # multi-channel EXR with the layer names of Denoiser::saveBuffers (denoiser.cpp:88-112)
    path = str(tmp_path / "features.exr")
    it.features_write(path)
    chans, img = _read_exr_channels(path)
    assert chans == ["albedo.B", "albedo.G", "albedo.R", "color.B", "color.G", "color.R", "normal.B", "normal.G", "normal.R"]
    np.testing.assert_array_equal(img[..., [5, 4, 3]], got[..., 0:3])
    np.testing.assert_array_equal(img[..., [2, 1, 0]], got[..., 3:6])
    np.testing.assert_array_equal(img[..., [8, 7, 6]], got[..., 6:9])
    it.film_clear()                                      # clears the feature sums as well
    assert it.features()[..., 9].sum() == 0
