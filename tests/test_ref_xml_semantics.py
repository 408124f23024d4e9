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


# ---------------------------------------------------------------------------------------------------------------------
#  transforms, sensor, shapes, object defaults
# ---------------------------------------------------------------------------------------------------------------------
def xf_xml(ops):
    """'/'-separated ops of ref_harness.cpp's parseProps -> the <transform> children the scene handler reads"""
    out = ""
    for op in ops.split("/"):
        v = op[2:].split(",")
        if op[0] == "t":
            out += '<translate x="%s" y="%s" z="%s"/>' % tuple(v)
        elif op[0] == "r":
            out += '<rotate x="%s" y="%s" z="%s" angle="%s"/>' % tuple(v)
        elif op[0] == "s":
            out += '<scale x="%s" y="%s" z="%s"/>' % tuple(v)
        elif op[0] == "l":
            out += '<lookat origin="%s" target="%s"%s/>' % (", ".join(v[:3]), ", ".join(v[3:6]), (' up="%s"' % ", ".join(v[6:9])) if len(v) >= 9 else "")
        elif op[0] == "m":
            out += '<matrix value="%s"/>' % " ".join(v)
    return out


def props_xml(props):
    out = ""
    for name, kind, value in props:
        if kind == "x":
            out += '<transform name="%s">%s</transform>' % (name, xf_xml(value))
        elif kind == "c":
            out += '<rgb name="%s" value="%s"/>' % (name, value)
        else:
            out += '<%s name="%s" value="%s"/>' % (XML_TAG[kind], name, value)
    return out


SENSORS = {
    "defaults": [],
    "lookat_fov": [("toWorld", "x", "l:0,1,3.9,0,1,0,0,1,0"), ("fov", "f", "39.3")],
    "lookat_no_up": [("toWorld", "x", "l:1,2,3,0,0.5,-1"), ("fov", "f", "55"), ("fovAxis", "s", "y")],
    "ops_in_order": [("toWorld", "x", "r:0,1,0,35/t:1,2,3/r:1,0,0,-10/s:1,1,1"), ("fov", "f", "70"), ("fovAxis", "s", "diagonal")],
    "matrix_and_clip": [("toWorld", "x", "m:0,0,1,2,0,1,0,1,-1,0,0,3,0,0,0,1"), ("fov", "f", "30"), ("fovAxis", "s", "smaller"),
                        ("nearClip", "f", "0.5"), ("farClip", "f", "50")],
    "focal_length": [("focalLength", "s", "35mm"), ("fovAxis", "s", "larger")],
}


@pytest.mark.parametrize("case", sorted(SENSORS))
def test_sensor_properties_and_transform_composition(pkg, oracle, tmp_path, case):
    from b200pg import api

    props = SENSORS[case]
    W, H = 96, 40
    xml = ('<scene version="0.6.0"><integrator type="progressivepath"/><sensor type="perspective">%s<sampler type="independent"/>'
           '<film type="hdrfilm"><integer name="width" value="%d"/><integer name="height" value="%d"/></film></sensor>'
           '<shape type="rectangle"><emitter type="area"><rgb name="radiance" value="1"/></emitter></shape></scene>') % (props_xml(props), W, H)
    path = tmp_path / "s.xml"
    path.write_text(xml)
    sc = api.Scene.load_xml(str(path))
    osc = OracleScene.from_desc(oracle, sc.desc, keep=sc)
    pos = (np.random.RandomState(3).rand(1500, 2) * [W, H]).astype(np.float32)
    np.testing.assert_allclose(osc.camera_rays(pos), ref_lib.sensor_rays_from_props(props, W, H, pos), rtol=2e-6, atol=3e-6)


SHAPES = {
    "rectangle_default": ("rectangle", []),
    "rectangle_ops": ("rectangle", [("toWorld", "x", "s:2,0.5,1/r:1,0,0,-90/r:0,1,0,30/t:0.5,-1,1")]),
    "rectangle_flipped": ("rectangle", [("toWorld", "x", "r:0,1,0,120/t:0,0.3,0"), ("flipNormals", "b", "true")]),
    "cube_default": ("cube", []),
    "cube_ops": ("cube", [("toWorld", "x", "s:0.3,0.6,0.2/r:0,1,0,17/r:1,0,0,5/t:-0.4,0.55,-0.35")]),
    "cube_flipped": ("cube", [("toWorld", "x", "s:0.5,0.5,0.5"), ("flipNormals", "b", "true")]),
}


@pytest.mark.parametrize("case", sorted(SHAPES))
def test_shape_plugins_and_their_transforms(pkg, oracle, tmp_path, case):
    """rectangle.cpp / cube.cpp with toWorld built from XML ops and flipNormals: hit distance, position, geometric and shading
    normal, the shading frame's tangent (per-triangle UV tangents on the cube) of chords through the shape."""
    from b200pg import api

    plugin, props = SHAPES[case]
    xml = ('<scene version="0.6.0"><integrator type="progressivepath"/><sensor type="perspective"><sampler type="independent"/>'
           '<film type="hdrfilm"><integer name="width" value="8"/><integer name="height" value="8"/></film></sensor>'
           '<shape type="%s">%s<bsdf type="diffuse"/></shape>'
           '<shape type="rectangle"><transform name="toWorld"><translate x="50" y="50" z="50"/></transform>'
           '<emitter type="area"><rgb name="radiance" value="1"/></emitter></shape></scene>') % (plugin, props_xml(props))
    path = tmp_path / "s.xml"
    path.write_text(xml)
    sc = api.Scene.load_xml(str(path))
    osc = OracleScene.from_desc(oracle, sc.desc, keep=sc)
    rng = np.random.RandomState(8)
    a, b = rng.randn(6000, 3), rng.randn(6000, 3) * 0.4
    a = 4.0 * a / np.linalg.norm(a, axis=1, keepdims=True)
    d = b - a
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([a, np.zeros((6000, 1)), d, np.full((6000, 1), np.inf)], 1).astype(np.float32)
    o, r = osc.intersect(rays), ref_lib.shape_hits_from_props(plugin, props, rays)
    hit = np.isfinite(r["t"])
    assert np.array_equal(np.isfinite(o["t"]), hit) and hit.mean() > 0.03
    np.testing.assert_allclose(o["t"][hit], r["t"][hit], rtol=1e-5, atol=1e-5)
    for k in ("p", "geo_n", "sh_n", "sh_s"):
        assert np.abs(o[k][hit] - r[k][hit]).max() <= 3e-5, k


def test_object_defaults(pkg, tmp_path):
    """film.cpp:29-32 (768 x 576), the gaussian filter's radius (gaussian.cpp:33-37), independent.cpp's sampleCount, the area
    light's samplingWeight -- as the reference's constructors set them without any property, against the reader's."""
    from b200pg import api

    path = tmp_path / "s.xml"
    path.write_text('<scene version="0.6.0"><integrator type="progressivepath"/><sensor type="perspective"><sampler type="independent"/>'
                    '<film type="hdrfilm"/></sensor><shape type="rectangle"><emitter type="area"><rgb name="radiance" value="1"/>'
                    '</emitter></shape></scene>')
    sc = api.Scene.load_xml(str(path))
    d, r = sc.desc, ref_lib.defaults()
    assert (d.film.width, d.film.height) == (r["film_width"], r["film_height"])
    assert abs(4 * d.film.filter_stddev - r["filter_radius"]) < 1e-6
    assert d.sample_count == r["sample_count"]
    assert d.emitters[0].sampling_weight == r["sampling_weight"]


def test_fallback_camera_of_a_scene_without_a_sensor(pkg, oracle, tmp_path):
    """Scene::configure (scene.cpp:272-305) run by the reference on the same shapes: the reader's fallback camera must produce
    the rays of the camera the reference adds, on the film the reference defaults to."""
    from b200pg import api

    path = tmp_path / "min.xml"
    path.write_text("""<scene version="0.6.0">
      <shape type="rectangle"><transform name="toWorld"><scale x="2" y="1"/><translate x="1" y="3" z="5"/></transform></shape>
      <shape type="cube"><transform name="toWorld"><scale value="0.5"/><translate x="0" y="3" z="8"/></transform>
        <emitter type="area"><rgb name="radiance" value="2"/></emitter></shape>
    </scene>""")
    sc = api.Scene.load_xml(str(path))
    osc = OracleScene.from_desc(oracle, sc.desc, keep=sc)
    rs = ref_lib.RefScene(desc=sc.desc, keep=sc, without_sensor=True)
    assert (rs.W, rs.H) == (sc.desc.film.width, sc.desc.film.height) == (768, 576)
    pos = (np.random.RandomState(2).rand(2000, 2) * [768, 576]).astype(np.float32)
    np.testing.assert_allclose(osc.camera_rays(pos), rs.camera_rays(pos), rtol=3e-6, atol=3e-6)


MEDIA = {
    # (medium properties, extra gridvolume properties, phase plugin, phase properties)
    "defaults": ([], [], "", []),                                                     # woodcock, scale 1, isotropic phase
    "woodcock_hg": ([("method", "s", "woodcock"), ("scale", "f", "20")], [], "hg", [("g", "f", "0.7")]),
    "hg_default_g": ([("scale", "f", "8")], [], "hg", []),
    "simpson_stepsize": ([("method", "s", "simpson"), ("scale", "f", "12"), ("stepSize", "f", "0.02")], [], "isotropic", []),
}


@pytest.mark.parametrize("case", sorted(MEDIA))
def test_heterogeneous_medium_from_xml_and_the_vol_file(pkg, oracle, tmp_path, case):
    """<medium type="heterogeneous"> with a gridvolume read from a .vol file (gridvolume.cpp:218-290), through the product's
    XML reader and through the reference's own plugins: density look-ups, free-flight distances (same stream), transmittance,
    phase function -- defaults, both tracking methods, an explicit stepSize, a volume-to-world transform."""
    from b200pg import api

    mprops, dprops, phase_plugin, pprops = MEDIA[case]
    S = pkg.scenes
    dens = S.fbm_density(res=24, seed=7)
    vol = tmp_path / "d.vol"
    S.write_vol(str(vol), dens, (-0.5, 0.0, -0.5), (0.5, 1.0, 0.5))
    phase_xml = ('<phase type="%s">%s</phase>' % (phase_plugin, props_xml(pprops))) if phase_plugin else ""
    xml = ('<scene version="0.6.0"><integrator type="progressivevolpath"/>'
           '<medium type="heterogeneous" id="m">%s<volume name="density" type="gridvolume"><string name="filename" value="%s"/>%s</volume>'
           '<volume name="albedo" type="constvolume"><spectrum name="value" value="0.9"/></volume>%s</medium>'
           '<sensor type="perspective"><sampler type="independent"/><film type="hdrfilm"><integer name="width" value="8"/>'
           '<integer name="height" value="8"/></film></sensor>'
           '<shape type="cube"><transform name="toWorld"><scale value="3"/></transform><ref name="interior" id="m"/></shape>'
           '<shape type="rectangle"><transform name="toWorld"><translate x="0" y="0" z="9"/></transform>'
           '<emitter type="area"><rgb name="radiance" value="1"/></emitter></shape></scene>') % (
               props_xml(mprops), vol, props_xml(dprops), phase_xml)
    path = tmp_path / "s.xml"
    path.write_text(xml)
    sc = api.Scene.load_xml(str(path))
    osc = OracleScene.from_desc(oracle, sc.desc, keep=sc)
    rm = ref_lib.medium_from_props(mprops, [("filename", "s", str(vol))] + dprops, [("value", "c", "0.9,0.9,0.9")], phase_plugin, pprops,
                                   sc.desc.seed)
    rng = np.random.RandomState(4)
    pts = (rng.rand(4000, 3) * [2.0, 1.6, 2.0] - [1.0, 0.3, 1.0]).astype(np.float32)
    np.testing.assert_allclose(osc.grid_lookup(0, pts), rm.grid_lookup(0, pts), rtol=1e-5, atol=2e-6)
    # segments strictly inside the grid's box (a start ON the boundary is a knife edge in the reference itself)
    lo, hi = np.array(sc.desc.media[0].aabb_min[:]), np.array(sc.desc.media[0].aabb_max[:])
    a = lo + (hi - lo) * (0.05 + 0.9 * rng.rand(2000, 3))
    b = lo + (hi - lo) * (0.05 + 0.9 * rng.rand(2000, 3))
    d = b - a
    L = np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([a, np.zeros((2000, 1)), d / L, L * (0.3 + rng.rand(2000, 1))], 1).astype(np.float32)
    rays[:, 7] = np.minimum(rays[:, 7], (L[:, 0] * 0.999).astype(np.float32))
    to, tro, _, _ = osc.medium_sample(0, rays)
    tr, _, trr = rm.medium_sample(0, rays)
    assert np.array_equal(np.isfinite(to), np.isfinite(tr)) and 0.05 < np.isfinite(tr).mean() < 0.98
    m = np.isfinite(tr)
    np.testing.assert_allclose(to[m], tr[m], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(tro, trr, rtol=2e-5, atol=2e-6)
    wi, wo, u = random_dirs(rng, 1000), random_dirs(rng, 1000), rng.rand(1000, 2).astype(np.float32)
    eo, er = osc.phase(0, wi, wo, u), rm.phase(0, wi, wo, u)
    np.testing.assert_allclose(eo[0], er[0], rtol=3e-5)
    np.testing.assert_allclose(eo[1], er[1], atol=1e-4)  # strongly forward hg: the inversion amplifies rounding


def test_volume_transforms_are_refused_not_dropped(pkg, tmp_path):
    """gridvolume.cpp:110-117 lets `toWorld` / `min` / `max` place the grid. The device keeps an axis-aligned worldToGrid, so the
    reader refuses them by name (a silently ignored transform renders a different scene -- which this comparison showed it did)."""
    from b200pg import api

    S = pkg.scenes
    vol = tmp_path / "d.vol"
    S.write_vol(str(vol), S.fbm_density(res=8, seed=7), (-0.5, 0.0, -0.5), (0.5, 1.0, 0.5))
    xml = ('<scene version="0.6.0"><integrator type="progressivevolpath"/><medium type="heterogeneous" id="m">'
           '<volume name="density" type="gridvolume"><string name="filename" value="%s"/>%s</volume>'
           '<volume name="albedo" type="constvolume"><spectrum name="value" value="0.9"/></volume></medium>'
           '<shape type="cube"><ref name="interior" id="m"/></shape>'
           '<shape type="rectangle"><emitter type="area"><rgb name="radiance" value="1"/></emitter></shape></scene>')
    for extra, word in (('<transform name="toWorld"><scale x="1.5" y="0.8" z="1.2"/></transform>', "toWorld"),
                        ('<point name="min" x="0" y="0" z="0"/>', "min")):
        path = tmp_path / "s.xml"
        path.write_text(xml % (vol, extra))
        with pytest.raises(api.B200pgError, match=word):
            api.Scene.load_xml(str(path))
    path.write_text(xml % (vol, '<transform name="toWorld"><translate x="0" y="0" z="0"/></transform>'))  # identity is fine
    api.Scene.load_xml(str(path)).close()
