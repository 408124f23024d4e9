"""BSDF parameterisations of the reference's data/tests/test_bsdf.xml that are on the hot path
(diffuse :15, twosided(diffuse) :21-23, dielectric water/air :45-48, roughconductor beckmann a=.3 :115-118,
roughplastic beckmann a=.7 :131-134) plus the GGX / anisotropic variants the benchmark scenes use."""
import numpy as np

from conftest import load_package

b200pg = load_package()
S = b200pg.scenes

CU_ETA, CU_K = (0.2004, 0.9240, 1.1022), (3.9129, 2.4528, 2.1421)  # copper, linear RGB (explicit, no SPD integration)


def bsdf_scene():
    """One tiny scene holding every BSDF under test; returns (builder, {name: bsdf index})."""
    sb = S.SceneBuilder(16, 16, spp=1)
    idx = {}
    idx["diffuse"] = sb.diffuse((0.5, 0.5, 0.5))
    idx["twosided_diffuse"] = sb.diffuse((0.5, 0.5, 0.5), twosided=True)
    idx["dielectric_water_air"] = sb.dielectric(int_ior=S.IOR["water"], ext_ior=S.IOR["air"])
    idx["roughconductor_beckmann_0.3"] = sb.roughconductor(CU_ETA, CU_K, alpha=0.3, distribution="beckmann")
    idx["roughconductor_ggx_0.15"] = sb.roughconductor(CU_ETA, CU_K, alpha=0.15, distribution="ggx")
    idx["roughplastic_beckmann_0.7"] = sb.roughplastic(alpha=0.7, distribution="beckmann")
    idx["roughplastic_ggx_0.2_twosided"] = sb.roughplastic(diffuse=(0.4, 0.25, 0.1), alpha=0.2, distribution="ggx", twosided=True)
    sb.rectangle([S.scale(1, 1, 1)], bsdf=idx["diffuse"], radiance=(1, 1, 1))
    for i in range(1, len(idx)):
        sb.rectangle([S.translate(3.0 * i, 0, 0)], bsdf=i)
    sb.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0)
    return sb, idx


def random_dirs(rng, n, upper=False):
    d = rng.randn(n, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if upper:
        d[:, 2] = np.abs(d[:, 2])
    return d.astype(np.float32)
