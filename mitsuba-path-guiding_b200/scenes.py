"""Seeded synthetic scenes for the BASELINE configs, built twice:

* as a flat ``SceneDesc`` (include/b200pg.h) for ``b200pg_scene_from_arrays`` and the oracle, and
* as ordinary Mitsuba 0.6 XML text (``to_xml``) for ``b200pg_scene_load_xml`` -- the same XML real
  Mitsuba would load.

Transforms follow the reference: ops compose as ``op * current`` (scenehandler.cpp:348-440),
``rotate`` per transform.cpp:65-98, ``lookAt`` per transform.cpp:191-214. Everything is float32.
"""
import ctypes as C
import math

import numpy as np

from . import _abi as A

f32 = np.float32


# ----------------------------------------------------------------------------- transforms
def translate(x, y, z):
    m = np.eye(4, dtype=f32)
    m[0, 3], m[1, 3], m[2, 3] = x, y, z
    return m


def scale(x, y, z):
    m = np.eye(4, dtype=f32)
    m[0, 0], m[1, 1], m[2, 2] = x, y, z
    return m


def rotate(axis, angle_deg):
    a = np.asarray(axis, dtype=f32)
    a = a / f32(np.sqrt(np.dot(a, a)))
    th = f32(angle_deg) * f32(math.pi / 180.0)
    s, c = f32(np.sin(th)), f32(np.cos(th))
    x, y, z = a
    one = f32(1.0)
    m = np.eye(4, dtype=f32)
    m[0, 0] = x * x + (one - x * x) * c
    m[0, 1] = x * y * (one - c) - z * s
    m[0, 2] = x * z * (one - c) + y * s
    m[1, 0] = x * y * (one - c) + z * s
    m[1, 1] = y * y + (one - y * y) * c
    m[1, 2] = y * z * (one - c) - x * s
    m[2, 0] = x * z * (one - c) - y * s
    m[2, 1] = y * z * (one - c) + x * s
    m[2, 2] = z * z + (one - z * z) * c
    return m


def look_at(origin, target, up):
    p = np.asarray(origin, dtype=f32)
    t = np.asarray(target, dtype=f32)
    u = np.asarray(up, dtype=f32)
    d = t - p
    d = d / f32(np.sqrt(np.dot(d, d)))
    left = np.cross(u, d).astype(f32)
    left = left / f32(np.sqrt(np.dot(left, left)))
    new_up = np.cross(d, left).astype(f32)
    m = np.eye(4, dtype=f32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, p
    return m


def compose(*ops):
    """ops in XML order (first applied first): result = op_n * ... * op_1."""
    m = np.eye(4, dtype=f32)
    for op in ops:
        m = (op.astype(f32) @ m).astype(f32)
    return m


# Cube data of src/shapes/cube.cpp:24-30 (24 vertices, per-face normals/uvs, 12 triangles)
_CUBE_P = np.array(
    [[1, -1, -1], [1, -1, 1], [-1, -1, 1], [-1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, 1, 1], [1, 1, 1],
     [1, -1, -1], [1, 1, -1], [1, 1, 1], [1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1],
     [-1, -1, 1], [-1, 1, 1], [-1, 1, -1], [-1, -1, -1], [1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1]], dtype=f32)
_CUBE_N = np.array(
    [[0, -1, 0]] * 4 + [[0, 1, 0]] * 4 + [[1, 0, 0]] * 4 + [[0, 0, 1]] * 4 + [[-1, 0, 0]] * 4 + [[0, 0, -1]] * 4, dtype=f32)
_CUBE_UV = np.array([[0, 1], [1, 1], [1, 0], [0, 0]] * 6, dtype=f32)
_CUBE_T = np.array(
    [[0, 1, 2], [3, 0, 2], [4, 5, 6], [7, 4, 6], [8, 9, 10], [11, 8, 10], [12, 13, 14], [15, 12, 14],
     [16, 17, 18], [19, 16, 18], [20, 21, 22], [23, 20, 22]], dtype=np.uint32)


def _xform_point(m, p):
    q = (m[:3, :3] @ p.T).T + m[:3, 3]
    return np.ascontiguousarray(q, dtype=f32)


def _xform_normal(m, n):
    inv = np.linalg.inv(m.astype(np.float64))
    q = (inv[:3, :3].T @ n.astype(np.float64).T).T
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(q, dtype=f32)


IOR = {"vacuum": 1.0, "air": 1.000277, "water": 1.3330, "bk7": 1.5046, "polypropylene": 1.49,
       "acrylic glass": 1.49, "diamond": 2.419}  # src/bsdfs/ior.h:39-64 (subset)


class SceneBuilder:
    """Collects shapes/bsdfs/emitters/media; emits a ctypes SceneDesc and Mitsuba XML."""

    def __init__(self, width, height, spp=64, seed=1337):
        self.shapes, self.bsdfs, self.emitters, self.media = [], [], [], []
        self.width, self.height, self.spp, self.seed = width, height, spp, seed
        self.sensor = dict(to_world=np.eye(4, dtype=f32), fov=39.3, fov_axis=0, near=1e-2, far=1e4, medium=-1,
                           xml_lookat=None)
        self.filter_stddev = 0.5
        self.integrator = dict(type="progressivepath", maxDepth=8)
        self._keep = []

    # ---- bsdfs
    def _bsdf(self, **kw):
        d = dict(type=A.BSDF_DIFFUSE, twosided=0, reflectance=(0.5, 0.5, 0.5), specular_reflectance=(1, 1, 1),
                 specular_transmittance=(1, 1, 1), int_ior=IOR["bk7"], ext_ior=IOR["air"], eta=(0, 0, 0), k=(1, 1, 1),
                 distribution=A.DISTR_BECKMANN, alpha=0.1, nonlinear=0)
        d.update(kw)
        self.bsdfs.append(d)
        return len(self.bsdfs) - 1

    def diffuse(self, rgb, twosided=False):
        return self._bsdf(type=A.BSDF_DIFFUSE, reflectance=tuple(rgb), twosided=int(twosided))

    def dielectric(self, int_ior=IOR["bk7"], ext_ior=IOR["air"]):
        return self._bsdf(type=A.BSDF_DIELECTRIC, int_ior=int_ior, ext_ior=ext_ior)

    def roughconductor(self, eta, k, alpha=0.1, distribution="beckmann", twosided=False):
        return self._bsdf(type=A.BSDF_ROUGHCONDUCTOR, eta=tuple(eta), k=tuple(k), alpha=alpha, ext_ior=IOR["air"],
                          distribution=A.DISTR_GGX if distribution == "ggx" else A.DISTR_BECKMANN, twosided=int(twosided))

    def roughplastic(self, diffuse=(0.5, 0.5, 0.5), alpha=0.1, int_ior=IOR["polypropylene"], ext_ior=IOR["air"],
                     distribution="beckmann", twosided=False, nonlinear=False):
        return self._bsdf(type=A.BSDF_ROUGHPLASTIC, reflectance=tuple(diffuse), alpha=alpha, int_ior=int_ior,
                          ext_ior=ext_ior, distribution=A.DISTR_GGX if distribution == "ggx" else A.DISTR_BECKMANN,
                          twosided=int(twosided), nonlinear=int(nonlinear))

    # ---- shapes
    def rectangle(self, ops, bsdf=-1, radiance=None, interior=-1, exterior=-1, xml_ops=None):
        m = compose(*ops)
        s = dict(type=A.SHAPE_RECTANGLE, to_world=m, bsdf=bsdf, emitter=-1, interior=interior, exterior=exterior,
                 xml="rectangle", xml_ops=xml_ops)
        self.shapes.append(s)
        if radiance is not None:
            self.emitters.append(dict(radiance=tuple(radiance), weight=1.0, shape=len(self.shapes) - 1))
            s["emitter"] = len(self.emitters) - 1
        return len(self.shapes) - 1

    def cube(self, ops, bsdf=-1, radiance=None, interior=-1, exterior=-1, xml_ops=None):
        m = compose(*ops)
        s = dict(type=A.SHAPE_TRIMESH, to_world=np.eye(4, dtype=f32), bsdf=bsdf, emitter=-1, interior=interior,
                 exterior=exterior, P=_xform_point(m, _CUBE_P), N=_xform_normal(m, _CUBE_N), UV=_CUBE_UV.copy(),
                 T=_CUBE_T.copy(), xml="cube", xml_ops=xml_ops, xml_matrix=m)
        self.shapes.append(s)
        if radiance is not None:
            self.emitters.append(dict(radiance=tuple(radiance), weight=1.0, shape=len(self.shapes) - 1))
            s["emitter"] = len(self.emitters) - 1
        return len(self.shapes) - 1

    def trimesh(self, P, T, N=None, UV=None, bsdf=-1, radiance=None, name=None):
        s = dict(type=A.SHAPE_TRIMESH, to_world=np.eye(4, dtype=f32), bsdf=bsdf, emitter=-1, interior=-1, exterior=-1,
                 P=np.ascontiguousarray(P, dtype=f32), N=None if N is None else np.ascontiguousarray(N, dtype=f32),
                 UV=None if UV is None else np.ascontiguousarray(UV, dtype=f32),
                 T=np.ascontiguousarray(T, dtype=np.uint32), xml="serialized", xml_ops=None, name=name)
        self.shapes.append(s)
        if radiance is not None:
            self.emitters.append(dict(radiance=tuple(radiance), weight=1.0, shape=len(self.shapes) - 1))
            s["emitter"] = len(self.emitters) - 1
        return len(self.shapes) - 1

    def medium(self, density, aabb_min, aabb_max, scale_=1.0, albedo=(0.9, 0.9, 0.9), phase="isotropic", g=0.0,
               method="woodcock", vol_path=None):
        d = np.ascontiguousarray(density, dtype=f32)  # [z][y][x]
        self.media.append(dict(density=d, aabb_min=tuple(aabb_min), aabb_max=tuple(aabb_max), scale=scale_,
                               albedo=tuple(albedo), phase=A.PHASE_HG if phase == "hg" else A.PHASE_ISOTROPIC, g=g,
                               method=A.MEDIUM_SIMPSON if method == "simpson" else A.MEDIUM_WOODCOCK, vol_path=vol_path))
        return len(self.media) - 1

    def set_camera(self, origin, target, up, fov, fov_axis=0, medium=-1):
        self.sensor.update(to_world=look_at(origin, target, up), fov=fov, fov_axis=fov_axis, medium=medium,
                           xml_lookat=(origin, target, up))

    # ---- flat description
    def desc(self, rtrans_reduce=None):
        """Returns (SceneDesc, keepalive). ``rtrans_reduce(distribution, eta, alpha)`` must return
        (ext_trans[100], ext_diff, int_diff) for roughplastic entries (tables come from the product's
        loader or from the oracle's reduction, depending on who asks)."""
        keep = []
        shapes = (A.Shape * max(1, len(self.shapes)))()
        for i, s in enumerate(self.shapes):
            o = shapes[i]
            o.type = s["type"]
            o.to_world[:] = s["to_world"].astype(f32).ravel().tolist()
            o.bsdf, o.emitter = s["bsdf"], s["emitter"]
            o.interior_medium, o.exterior_medium = s["interior"], s["exterior"]
            if s["type"] == A.SHAPE_TRIMESH:
                o.n_vertices, o.n_triangles = s["P"].shape[0], s["T"].shape[0]
                o.positions = s["P"].ctypes.data_as(A.c_float_p)
                o.normals = s["N"].ctypes.data_as(A.c_float_p) if s.get("N") is not None else None
                o.texcoords = s["UV"].ctypes.data_as(A.c_float_p) if s.get("UV") is not None else None
                o.indices = s["T"].ctypes.data_as(A.c_u32_p)
                keep += [s["P"], s.get("N"), s.get("UV"), s["T"]]
        bsdfs = (A.Bsdf * max(1, len(self.bsdfs)))()
        for i, b in enumerate(self.bsdfs):
            o = bsdfs[i]
            o.type, o.twosided = b["type"], b["twosided"]
            o.reflectance[:] = b["reflectance"]
            o.specular_reflectance[:] = b["specular_reflectance"]
            o.specular_transmittance[:] = b["specular_transmittance"]
            o.int_ior, o.ext_ior = b["int_ior"], b["ext_ior"]
            ext = f32(b["ext_ior"]) if b["type"] == A.BSDF_ROUGHCONDUCTOR else f32(1.0)
            o.eta[:] = [float(f32(v) / ext) for v in b["eta"]]  # roughconductor.cpp:188-189
            o.k[:] = [float(f32(v) / ext) for v in b["k"]]
            o.distribution = b["distribution"]
            o.alpha_u = o.alpha_v = b["alpha"]
            o.sample_visible, o.nonlinear = 1, b["nonlinear"]
            if b["type"] == A.BSDF_ROUGHPLASTIC:
                if rtrans_reduce is None:
                    raise ValueError("roughplastic needs rtrans_reduce")
                eta = float(f32(b["int_ior"]) / f32(b["ext_ior"]))
                tr, ed, idf = rtrans_reduce(b["distribution"], eta, b["alpha"])
                o.rt_ext_trans[:] = [float(v) for v in tr]
                o.rt_ext_diff, o.rt_int_diff = float(ed), float(idf)
        emitters = (A.Emitter * max(1, len(self.emitters)))()
        for i, e in enumerate(self.emitters):
            emitters[i].radiance[:] = e["radiance"]
            emitters[i].sampling_weight = e["weight"]
            emitters[i].shape = e["shape"]
        media = (A.Medium * max(1, len(self.media)))()
        for i, m in enumerate(self.media):
            o = media[i]
            o.method, o.scale = m["method"], m["scale"]
            o.albedo[:] = m["albedo"]
            o.phase_type, o.phase_g = m["phase"], m["g"]
            nz, ny, nx = m["density"].shape
            o.res[:] = [nx, ny, nz]
            o.aabb_min[:] = m["aabb_min"]
            o.aabb_max[:] = m["aabb_max"]
            o.to_world[:] = np.eye(4, dtype=f32).ravel().tolist()
            o.density = m["density"].ctypes.data_as(A.c_float_p)
            o.step_size_multiplier = 0.0
            keep.append(m["density"])
        d = A.SceneDesc()
        d.n_shapes, d.n_bsdfs, d.n_emitters, d.n_media = len(self.shapes), len(self.bsdfs), len(self.emitters), len(self.media)
        d.shapes, d.bsdfs, d.emitters, d.media = shapes, bsdfs, emitters, media
        d.sensor.to_world[:] = self.sensor["to_world"].astype(f32).ravel().tolist()
        d.sensor.fov, d.sensor.fov_axis = self.sensor["fov"], self.sensor["fov_axis"]
        d.sensor.near_clip, d.sensor.far_clip = self.sensor["near"], self.sensor["far"]
        d.sensor.medium = self.sensor["medium"]
        d.film.width, d.film.height, d.film.filter_stddev = self.width, self.height, self.filter_stddev
        d.sample_count, d.seed = self.spp, self.seed
        keep += [shapes, bsdfs, emitters, media]
        return d, keep

    # ---- Mitsuba 0.6 XML
    def to_xml(self, mesh_dir=None):
        def rgb(v):
            return "%.9g, %.9g, %.9g" % tuple(v)

        def ops_xml(ops, ind):
            out = []
            for op in ops or []:
                kind = op[0]
                if kind == "translate":
                    out.append('%s<translate x="%.9g" y="%.9g" z="%.9g"/>' % ((ind,) + tuple(op[1:])))
                elif kind == "scale":
                    out.append('%s<scale x="%.9g" y="%.9g" z="%.9g"/>' % ((ind,) + tuple(op[1:])))
                elif kind == "rotate":
                    out.append('%s<rotate x="%.9g" y="%.9g" z="%.9g" angle="%.9g"/>' % ((ind,) + tuple(op[1]) + (op[2],)))
            return out

        L = ['<?xml version="1.0" encoding="utf-8"?>', '<scene version="0.6.0">']
        integ = self.integrator
        L.append('    <integrator type="%s">' % integ["type"])
        for k, v in integ.items():
            if k == "type":
                continue
            if isinstance(v, bool):
                L.append('        <boolean name="%s" value="%s"/>' % (k, "true" if v else "false"))
            elif isinstance(v, int):
                L.append('        <integer name="%s" value="%d"/>' % (k, v))
            else:
                L.append('        <float name="%s" value="%.9g"/>' % (k, v))
        L.append("    </integrator>")
        for i, m in enumerate(self.media):
            L.append('    <medium type="heterogeneous" id="medium%d">' % i)
            L.append('        <string name="method" value="%s"/>' % ("simpson" if m["method"] == A.MEDIUM_SIMPSON else "woodcock"))
            L.append('        <float name="scale" value="%.9g"/>' % m["scale"])
            L.append('        <volume name="density" type="gridvolume">')
            L.append('            <string name="filename" value="%s"/>' % (m["vol_path"] or ("medium%d.vol" % i)))
            L.append("        </volume>")
            L.append('        <volume name="albedo" type="constvolume">')
            L.append('            <spectrum name="value" value="%s"/>' % rgb(m["albedo"]))
            L.append("        </volume>")
            if m["phase"] == A.PHASE_HG:
                L.append('        <phase type="hg"><float name="g" value="%.9g"/></phase>' % m["g"])
            else:
                L.append('        <phase type="isotropic"/>')
            L.append("    </medium>")
        L.append('    <sensor type="perspective">')
        L.append('        <float name="fov" value="%.9g"/>' % self.sensor["fov"])
        L.append('        <string name="fovAxis" value="%s"/>' % ["x", "y", "diagonal", "smaller", "larger"][self.sensor["fov_axis"]])
        L.append('        <float name="nearClip" value="%.9g"/>' % self.sensor["near"])
        L.append('        <float name="farClip" value="%.9g"/>' % self.sensor["far"])
        o, t, u = self.sensor["xml_lookat"]
        L.append('        <transform name="toWorld">')
        L.append('            <lookat origin="%.9g, %.9g, %.9g" target="%.9g, %.9g, %.9g" up="%.9g, %.9g, %.9g"/>' % (tuple(o) + tuple(t) + tuple(u)))
        L.append("        </transform>")
        if self.sensor["medium"] >= 0:
            L.append('        <ref id="medium%d"/>' % self.sensor["medium"])
        L.append('        <sampler type="independent"><integer name="sampleCount" value="%d"/></sampler>' % self.spp)
        L.append('        <film type="hdrfilm">')
        L.append('            <integer name="width" value="%d"/>' % self.width)
        L.append('            <integer name="height" value="%d"/>' % self.height)
        L.append('            <boolean name="banner" value="false"/>')
        L.append('            <string name="componentFormat" value="float32"/>')
        L.append('            <string name="fileFormat" value="pfm"/>')
        L.append('            <rfilter type="gaussian"><float name="stddev" value="%.9g"/></rfilter>' % self.filter_stddev)
        L.append("        </film>")
        L.append("    </sensor>")
        for i, b in enumerate(self.bsdfs):
            ind = "    "
            if b["twosided"]:
                L.append('    <bsdf type="twosided" id="bsdf%d">' % i)
                ind = "        "
                L.append('%s<bsdf type="%s">' % (ind, ["diffuse", "dielectric", "roughconductor", "roughplastic", "null"][b["type"]]))
            else:
                L.append('%s<bsdf type="%s" id="bsdf%d">' % (ind, ["diffuse", "dielectric", "roughconductor", "roughplastic", "null"][b["type"]], i))
            ii = ind + "    "
            if b["type"] == A.BSDF_DIFFUSE:
                L.append('%s<rgb name="reflectance" value="%s"/>' % (ii, rgb(b["reflectance"])))
            elif b["type"] == A.BSDF_DIELECTRIC:
                L.append('%s<float name="intIOR" value="%.9g"/>' % (ii, b["int_ior"]))
                L.append('%s<float name="extIOR" value="%.9g"/>' % (ii, b["ext_ior"]))
            elif b["type"] == A.BSDF_ROUGHCONDUCTOR:
                L.append('%s<string name="distribution" value="%s"/>' % (ii, "ggx" if b["distribution"] == A.DISTR_GGX else "beckmann"))
                L.append('%s<float name="alpha" value="%.9g"/>' % (ii, b["alpha"]))
                L.append('%s<rgb name="eta" value="%s"/>' % (ii, rgb(b["eta"])))
                L.append('%s<rgb name="k" value="%s"/>' % (ii, rgb(b["k"])))
                L.append('%s<float name="extEta" value="%.9g"/>' % (ii, b["ext_ior"]))
            elif b["type"] == A.BSDF_ROUGHPLASTIC:
                L.append('%s<string name="distribution" value="%s"/>' % (ii, "ggx" if b["distribution"] == A.DISTR_GGX else "beckmann"))
                L.append('%s<float name="alpha" value="%.9g"/>' % (ii, b["alpha"]))
                L.append('%s<float name="intIOR" value="%.9g"/>' % (ii, b["int_ior"]))
                L.append('%s<float name="extIOR" value="%.9g"/>' % (ii, b["ext_ior"]))
                L.append('%s<rgb name="diffuseReflectance" value="%s"/>' % (ii, rgb(b["reflectance"])))
                if b["nonlinear"]:
                    L.append('%s<boolean name="nonlinear" value="true"/>' % ii)
            L.append("%s</bsdf>" % ind)
            if b["twosided"]:
                L.append("    </bsdf>")
        for i, s in enumerate(self.shapes):
            L.append('    <shape type="%s">' % s["xml"])
            if s["xml"] == "serialized":
                L.append('        <string name="filename" value="%s"/>' % ((mesh_dir + "/" if mesh_dir else "") + (s.get("name") or "mesh%d" % i) + ".serialized"))
            if s.get("xml_ops"):
                L.append('        <transform name="toWorld">')
                L += ops_xml(s["xml_ops"], "            ")
                L.append("        </transform>")
            elif s["xml"] != "serialized":
                m = s["to_world"] if s["xml"] == "rectangle" else s["xml_matrix"]
                L.append('        <transform name="toWorld"><matrix value="%s"/></transform>' % " ".join("%.9g" % v for v in m.ravel()))
            if s["bsdf"] >= 0:
                L.append('        <ref id="bsdf%d"/>' % s["bsdf"])
            if s["emitter"] >= 0:
                L.append('        <emitter type="area"><rgb name="radiance" value="%s"/></emitter>' % rgb(self.emitters[s["emitter"]]["radiance"]))
            if s["interior"] >= 0:
                L.append('        <ref name="interior" id="medium%d"/>' % s["interior"])
            if s["exterior"] >= 0:
                L.append('        <ref name="exterior" id="medium%d"/>' % s["exterior"])
            L.append("    </shape>")
        L.append("</scene>")
        return "\n".join(L) + "\n"


# ----------------------------------------------------------------------------- named scenes
def _op(kind, *a):
    return (kind,) + a


def _cornell_walls(sb, with_boxes=True):
    white = sb.diffuse((0.725, 0.71, 0.68), twosided=True)
    red = sb.diffuse((0.63, 0.065, 0.05), twosided=True)
    green = sb.diffuse((0.14, 0.45, 0.091), twosided=True)

    def rect(xml_ops, bsdf, **kw):
        ops = []
        for o in xml_ops:
            if o[0] == "translate":
                ops.append(translate(*o[1:]))
            elif o[0] == "scale":
                ops.append(scale(*o[1:]))
            else:
                ops.append(rotate(o[1], o[2]))
        return ops

    X, Y, Z = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    walls = [
        ([_op("rotate", X, -90.0)], white),                                   # floor  y=0, n=+y
        ([_op("rotate", X, 90.0), _op("translate", 0.0, 2.0, 0.0)], white),   # ceiling y=2, n=-y
        ([_op("translate", 0.0, 1.0, -1.0)], white),                          # back z=-1, n=+z
        ([_op("rotate", Y, 90.0), _op("translate", -1.0, 1.0, 0.0)], red),    # left x=-1, n=+x
        ([_op("rotate", Y, -90.0), _op("translate", 1.0, 1.0, 0.0)], green),  # right x=1, n=-x
    ]
    for xo, b in walls:
        sb.rectangle(rect(xo, b), bsdf=b, xml_ops=xo)
    boxw = sb.diffuse((0.725, 0.71, 0.68))
    if with_boxes:
        # boxes float 1 mm above the floor: no coplanar faces, hence no equal-t ties whose winner would
        # depend on the traversal order of the accelerator (kd-tree in the reference, BVH here)
        tall = [_op("scale", 0.3, 0.6, 0.3), _op("rotate", Y, 17.0), _op("translate", -0.33, 0.601, -0.3)]
        short = [_op("scale", 0.3, 0.3, 0.3), _op("rotate", Y, -17.0), _op("translate", 0.35, 0.301, 0.35)]
        sb.cube(rect(tall, boxw), bsdf=boxw, xml_ops=tall)
        sb.cube(rect(short, boxw), bsdf=boxw, xml_ops=short)
    return rect, white, boxw


def cornell_box(width=512, height=512, spp=64, seed=1337, max_depth=8):
    """Config C1: procedural Cornell box from rectangle/cube shapes."""
    sb = SceneBuilder(width, height, spp, seed)
    rect, white, boxw = _cornell_walls(sb)
    X = (1, 0, 0)
    light = [_op("scale", 0.235, 0.19, 1.0), _op("rotate", X, 90.0), _op("translate", 0.0, 1.98, 0.0)]
    sb.rectangle(rect(light, -1), bsdf=-1, radiance=(17.0, 12.0, 4.0), xml_ops=light)
    sb.set_camera((0.0, 1.0, 3.9), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0), 39.3)
    sb.integrator = dict(type="progressivepath", maxDepth=max_depth)
    return sb


def cornell_caustic(width=1024, height=1024, spp=64, seed=1337, max_depth=8):
    """Config C2: small shielded light + glass cube -> mostly indirect / caustic transport."""
    sb = SceneBuilder(width, height, spp, seed)
    rect, white, boxw = _cornell_walls(sb, with_boxes=False)
    X, Y = (1, 0, 0), (0, 1, 0)
    s = 0.47 * 0.38 / (0.05 * 0.05)
    # small light facing UP towards the ceiling (shielded from below by a reflector plate)
    light = [_op("scale", 0.025, 0.025, 1.0), _op("rotate", X, -90.0), _op("translate", 0.0, 1.6, 0.0)]
    sb.rectangle(rect(light, -1), bsdf=-1, radiance=(17.0 * s, 12.0 * s, 4.0 * s), xml_ops=light)
    shield = [_op("scale", 0.15, 0.15, 1.0), _op("rotate", X, 90.0), _op("translate", 0.0, 1.55, 0.0)]
    sb.rectangle(rect(shield, white), bsdf=white, xml_ops=shield)
    glass = sb.dielectric(int_ior=1.5, ext_ior=1.0)
    # lifted 2 mm off the floor: coplanar glass/floor faces would make the closest hit ambiguous (equal t)
    gl = [_op("scale", 0.3, 0.3, 0.3), _op("rotate", Y, 25.0), _op("translate", 0.3, 0.302, 0.2)]
    sb.cube(rect(gl, glass), bsdf=glass, xml_ops=gl)
    tall = [_op("scale", 0.25, 0.55, 0.25), _op("rotate", Y, 17.0), _op("translate", -0.4, 0.551, -0.35)]
    sb.cube(rect(tall, boxw), bsdf=boxw, xml_ops=tall)
    sb.set_camera((0.0, 1.0, 3.9), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0), 39.3)
    sb.integrator = dict(type="progressivepath", maxDepth=max_depth)
    return sb


def fbm_density(res=64, seed=1337, octaves=3):
    """clamp(fBm(value noise), 0, 1) on a res^3 grid, array order [z][y][x]."""
    rng = np.random.RandomState(seed)
    out = np.zeros((res, res, res), dtype=np.float64)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        n = 4 * (2 ** o) + 1
        lattice = rng.rand(n, n, n)
        t = np.linspace(0, n - 1, res)
        i0 = np.minimum(t.astype(int), n - 2)
        f = t - i0
        f = f * f * (3 - 2 * f)

        def lerp_axis(a, axis):
            a0 = np.take(a, i0, axis=axis)
            a1 = np.take(a, i0 + 1, axis=axis)
            shape = [1, 1, 1]
            shape[axis] = res
            ff = f.reshape(shape)
            return a0 * (1 - ff) + a1 * ff

        v = lerp_axis(lerp_axis(lerp_axis(lattice, 0), 1), 2)
        out += amp * v
        tot += amp
        amp *= 0.5
    out = out / tot
    out = np.clip((out - 0.35) * 2.2, 0.0, 1.0)
    return out.astype(f32)


def cornell_medium(width=1024, height=1024, spp=64, seed=1337, max_depth=8, res=256, phase="hg", g=0.7, scale_=20.0):
    """Config C3: Cornell walls + an index-matched cube boundary (no BSDF) holding a heterogeneous medium."""
    sb = SceneBuilder(width, height, spp, seed)
    rect, white, boxw = _cornell_walls(sb, with_boxes=False)
    X = (1, 0, 0)
    light = [_op("scale", 0.235, 0.19, 1.0), _op("rotate", X, 90.0), _op("translate", 0.0, 1.98, 0.0)]
    sb.rectangle(rect(light, -1), bsdf=-1, radiance=(17.0, 12.0, 4.0), xml_ops=light)
    dens = fbm_density(res, seed)
    med = sb.medium(dens, (-0.6, 0.2, -0.6), (0.6, 1.4, 0.6), scale_=scale_, albedo=(0.9, 0.9, 0.9), phase=phase, g=g)
    box = [_op("scale", 0.6, 0.6, 0.6), _op("translate", 0.0, 0.8, 0.0)]
    sb.cube(rect(box, -1), bsdf=-1, interior=med, xml_ops=box)
    sb.set_camera((0.0, 1.0, 3.9), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0), 39.3)
    sb.integrator = dict(type="progressivevolpath", maxDepth=max_depth)
    return sb


def heightfield_mesh(n=2237, seed=1337, amp=0.05):
    """(n-1)^2*2 triangles over [-1,1]^2 in xz, y = amp * sum of seeded sines, with vertex normals."""
    rng = np.random.RandomState(seed)
    x = np.linspace(-1, 1, n, dtype=np.float64)
    X, Z = np.meshgrid(x, x, indexing="xy")
    Y = np.zeros_like(X)
    dYdx = np.zeros_like(X)
    dYdz = np.zeros_like(X)
    for _ in range(6):
        fx, fz = rng.uniform(2, 14, 2)
        ph = rng.uniform(0, 2 * np.pi)
        a = rng.uniform(0.3, 1.0)
        arg = fx * X + fz * Z + ph
        Y += a * np.sin(arg)
        dYdx += a * fx * np.cos(arg)
        dYdz += a * fz * np.cos(arg)
    Y *= amp / 3.0
    dYdx *= amp / 3.0
    dYdz *= amp / 3.0
    P = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(f32)
    N = np.stack([-dYdx, np.ones_like(X), -dYdz], -1).reshape(-1, 3)
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    idx = np.arange(n * n, dtype=np.uint32).reshape(n, n)
    a, b, c, d = idx[:-1, :-1], idx[:-1, 1:], idx[1:, :-1], idx[1:, 1:]
    T = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([b, c, d], -1).reshape(-1, 3)], 0)
    return P, N.astype(f32), T.astype(np.uint32)


def mesh_scene(width=2048, height=2048, spp=16, seed=1337, n=2237, max_depth=8):
    """Config C4 (n = 2237 -> 2 * 2236^2 = 10.0 M triangles): ~10M-triangle procedural mesh in 4 shapes with roughconductor/roughplastic, 2 area lights."""
    sb = SceneBuilder(width, height, spp, seed)
    cond = sb.roughconductor(eta=(0.2004, 0.9240, 1.1022), k=(3.9129, 2.4528, 2.1421), alpha=0.15, distribution="ggx")
    plast = sb.roughplastic(diffuse=(0.4, 0.25, 0.1), alpha=0.2, distribution="beckmann")
    P, N, T = heightfield_mesh(n, seed)
    # split triangles into 4 shapes by quadrant of the first vertex; alternate materials
    cx = P[T[:, 0], 0] >= 0
    cz = P[T[:, 0], 2] >= 0
    for q, (mx, mz) in enumerate([(False, False), (True, False), (False, True), (True, True)]):
        sel = (cx == mx) & (cz == mz)
        Tq = T[sel]
        used, inv = np.unique(Tq.ravel(), return_inverse=True)
        sb.trimesh(P[used], inv.reshape(-1, 3).astype(np.uint32), N=N[used], bsdf=cond if q in (0, 3) else plast,
                   name="quad%d" % q)
    X = (1, 0, 0)
    for lx, rad in ((-0.5, (30.0, 26.0, 20.0)), (0.55, (12.0, 18.0, 30.0))):
        ops = [_op("scale", 0.12, 0.12, 1.0), _op("rotate", X, 90.0), _op("translate", lx, 0.9, 0.1)]
        mats = []
        for o in ops:
            mats.append(translate(*o[1:]) if o[0] == "translate" else scale(*o[1:]) if o[0] == "scale" else rotate(o[1], o[2]))
        sb.rectangle(mats, bsdf=-1, radiance=rad, xml_ops=ops)
    sb.set_camera((0.0, 1.3, 2.4), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0)
    sb.integrator = dict(type="progressivepath", maxDepth=max_depth)
    return sb


def write_serialized(path, P, T, N=None, UV=None, name="mesh"):
    """Mitsuba .serialized v4 file (trimesh.cpp:175-270): 0x041C, version, zlib{flags, name, counts, arrays}, offset dictionary."""
    import struct
    import zlib

    flags = 0x1000  # single precision
    if N is not None:
        flags |= 0x0001
    if UV is not None:
        flags |= 0x0002
    body = struct.pack("<I", flags) + name.encode() + b"\0" + struct.pack("<QQ", P.shape[0], T.shape[0])
    body += np.ascontiguousarray(P, "<f4").tobytes()
    if N is not None:
        body += np.ascontiguousarray(N, "<f4").tobytes()
    if UV is not None:
        body += np.ascontiguousarray(UV, "<f4").tobytes()
    body += np.ascontiguousarray(T, "<u4").tobytes()
    with open(path, "wb") as f:
        f.write(struct.pack("<HH", 0x041C, 4))
        f.write(zlib.compress(body, 1))
        f.write(struct.pack("<QI", 0, 1))


def save_scene(sb, directory, name="scene.xml"):
    """Writes the XML plus the mesh / volume files it references; returns the XML path."""
    import os

    os.makedirs(directory, exist_ok=True)
    for i, s in enumerate(sb.shapes):
        if s["xml"] == "serialized":
            write_serialized(os.path.join(directory, (s.get("name") or "mesh%d" % i) + ".serialized"), s["P"], s["T"], s.get("N"), s.get("UV"))
    for i, m in enumerate(sb.media):
        write_vol(os.path.join(directory, m["vol_path"] or ("medium%d.vol" % i)), m["density"], m["aabb_min"], m["aabb_max"])
    path = os.path.join(directory, name)
    with open(path, "w") as f:
        f.write(sb.to_xml())
    return path


def write_vol(path, density, aabb_min, aabb_max):
    """Mitsuba VOL v3 file (gridvolume.cpp:56-89, 224-286): 'VOL',3, type=1 (f32), nx,ny,nz, channels, 6xf32 AABB."""
    nz, ny, nx = density.shape
    with open(path, "wb") as f:
        f.write(b"VOL\x03")
        f.write(np.array([1, nx, ny, nz, 1], dtype="<i4").tobytes())
        f.write(np.array(list(aabb_min) + list(aabb_max), dtype="<f4").tobytes())
        f.write(np.ascontiguousarray(density, dtype="<f4").tobytes())
