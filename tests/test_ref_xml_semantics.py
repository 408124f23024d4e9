"""The XML reader's BSDF semantics against the reference's own plugins. The reference's SceneHandler needs Xerces-C and is not in
the oracle/_ref build, but what it does with a <bsdf> element is mechanical -- element name -> typed Properties entry -- and
the meaning of a property (defaults, named indices of refraction, `alpha` vs `alphaU` / `alphaV`, distribution names,
`nonlinear`, the `twosided` adapter) lives in the PLUGIN constructors, which are compiled here. So: the same (plugin, named
properties) pair goes (a) as XML through the product's reader -> flat description -> oracle port BSDF, and (b) as Properties into
the reference's plugin; eval / pdf / sample of the two must agree."""
import numpy as np
import pytest

import ref_lib
from bsdf_cases import random_dirs
from oracle_lib import OracleScene

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built (no /root/reference here)")

XML_TAG = {"f": "float", "i": "integer", "b": "boolean", "s": "string"}

CASES = {
    "diffuse_default": ("diffuse", [], False),
    "diffuse_rgb_twosided": ("diffuse", [("reflectance", "c", "0.2,0.5,0.7")], True),
    "dielectric_default": ("dielectric", [], False),                                   # bk7 / air
    "dielectric_named": ("dielectric", [("intIOR", "s", "water"), ("extIOR", "s", "air")], False),
    "dielectric_numbers_tinted": ("dielectric", [("intIOR", "f", "1.7"), ("extIOR", "f", "1.1"), ("specularReflectance", "c", "0.9,0.8,0.7"),
                                                 ("specularTransmittance", "c", "0.6,0.7,0.8")], False),
    "roughconductor_alpha": ("roughconductor", [("material", "s", "none"), ("eta", "c", "0.2,0.9,1.1"), ("k", "c", "3.9,2.4,2.1"),
                                                ("alpha", "f", "0.25")], False),      # default distribution: beckmann
    "roughconductor_ggx_aniso_exteta": ("roughconductor", [("material", "s", "none"), ("eta", "c", "0.2,0.9,1.1"), ("k", "c", "3.9,2.4,2.1"),
                                                           ("distribution", "s", "ggx"), ("alphaU", "f", "0.1"), ("alphaV", "f", "0.35"),
                                                           ("extEta", "s", "water"), ("specularReflectance", "c", "0.9,0.9,0.5")], False),
    "roughconductor_none_default_eta": ("roughconductor", [("material", "s", "none")], True),   # eta 0, k 1: a perfect mirror lobe shape
    "roughplastic_default": ("roughplastic", [], False),                               # beckmann .1, polypropylene / air, diffuse .5
    "roughplastic_ggx_nonlinear": ("roughplastic", [("distribution", "s", "ggx"), ("alpha", "f", "0.3"), ("intIOR", "s", "acrylic glass"),
                                                    ("diffuseReflectance", "c", "0.4,0.25,0.1"), ("nonlinear", "b", "true"),
                                                    ("specularReflectance", "c", "0.8,0.9,1.0")], True),
}

SCENE = """<scene version="0.6.0"><integrator type="progressivepath"/>
<sensor type="perspective"><sampler type="independent"/><film type="hdrfilm"><integer name="width" value="8"/><integer name="height" value="8"/></film></sensor>
<shape type="rectangle">%s<emitter type="area"><rgb name="radiance" value="1"/></emitter></shape></scene>"""


def bsdf_xml(plugin, props, twosided):
    inner = '<bsdf type="%s">' % plugin
    for name, kind, value in props:
        inner += ('<rgb name="%s" value="%s"/>' % (name, value)) if kind == "c" else ('<%s name="%s" value="%s"/>' % (XML_TAG[kind], name, value))
    inner += "</bsdf>"
    return '<bsdf type="twosided">%s</bsdf>' % inner if twosided else inner


@pytest.mark.parametrize("case", sorted(CASES))
def test_bsdf_properties_mean_what_the_reference_plugins_say(pkg, oracle, tmp_path, case):
    from b200pg import api

    plugin, props, twosided = CASES[case]
    path = tmp_path / "s.xml"
    path.write_text(SCENE % bsdf_xml(plugin, props, twosided))
    sc = api.Scene.load_xml(str(path))
    osc = OracleScene.from_desc(oracle, sc.desc, keep=sc)
    rng = np.random.RandomState(5)
    n = 4000
    wi, wo, u = random_dirs(rng, n), random_dirs(rng, n), rng.rand(n, 2).astype(np.float32)
    index = sc.desc.shapes[0].bsdf
    o = osc.bsdf(index, wi, wo, u)
    r = ref_lib.bsdf_from_props(plugin, props, wi, wo, u, twosided=twosided)
    for key in ("eval", "pdf"):
        assert np.all(np.abs(o[key] - r[key]) <= 2e-5 * np.maximum(np.abs(r[key]), 1e-3)), key
    good = r["spdf"] > 0
    assert np.array_equal(o["spdf"] > 0, good)
    assert np.array_equal(o["flags"][good], r["flags"][good])
    assert np.abs(o["wo"][good] - r["wo"][good]).max() <= 2e-3
    w = (np.abs(o["weight"][good] - r["weight"][good]) / np.maximum(np.abs(r["weight"][good]), 1e-3)).max(1)
    assert np.quantile(w, 0.99) <= 2e-4 and (w > 2e-2).mean() < 5e-3
    assert good.mean() > 0.3
