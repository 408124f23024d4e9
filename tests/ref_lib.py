"""ctypes wrapper around oracle/_ref/libref_harness.so (TEST INFRASTRUCTURE): the REFERENCE's own classes, compiled by
oracle/Makefile.ref from the sources under /root/reference, behind the C-ABI of oracle/ref_harness/ref_harness.cpp.

The method names and array layouts follow tests/oracle_lib.py's OracleScene so that a test reads
`ref.intersect(rays)` next to `orc.intersect(rays)`. `available()` is False when the library has not been built (no
/root/reference at build time and no prebuilt oracle/_ref shipped): tests that need it skip, the committed fixtures under
tests/golden/upstream_*.npz (tests/golden/make_upstream.py) keep the pin without it."""
import ctypes as C
import os

import numpy as np

from conftest import ROOT, load_package

b200pg = load_package()
A = b200pg._abi

fp = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)

SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")


def available():
    return os.path.exists(SO)


def _f(a):
    return a.ctypes.data_as(fp)


def _u(a):
    return a.ctypes.data_as(u32p)


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = L = C.CDLL(SO, mode=C.RTLD_GLOBAL)
        L.ref_last_error.restype = C.c_char_p
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.POINTER(A.SceneDesc)]
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_kd_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ref_intersect.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_int, fp, u32p, C.c_int]
        L.ref_camera_rays.argtypes = [C.c_void_p, fp, C.c_size_t, fp]
        L.ref_bsdf.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_size_t, fp, fp, fp, fp, fp, u32p]
        L.ref_emitter_direct.argtypes = [C.c_void_p, fp, fp, fp, fp, C.c_size_t, fp, fp, fp]
        L.ref_radiance.argtypes = [C.c_void_p, C.POINTER(A.IntegratorParams), u32p, u32p, C.c_size_t, fp, fp]
        L.ref_film_splat.argtypes = [C.c_void_p, fp, fp, C.c_size_t, fp]
        L.ref_render.argtypes = [C.c_void_p, C.POINTER(A.IntegratorParams), C.c_int, C.c_int, fp, C.c_int, C.c_int,
                                 C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
        L.ref_render_plugin.argtypes = [C.c_void_p, C.POINTER(A.IntegratorParams), C.c_char_p, C.c_char_p, C.c_int, fp,
                                        C.POINTER(C.c_double)]
        L.ref_plugin_put_film.argtypes = [C.c_void_p, fp, fp]
        L.ref_grid_lookup.argtypes = [C.c_void_p, C.c_int, fp, C.c_size_t, fp]
        L.ref_medium_sample.argtypes = [C.c_void_p, C.c_int, fp, C.c_size_t, fp, fp, fp]
        L.ref_phase.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_size_t, fp, fp, fp]
        L.ref_num_threads.restype = C.c_int
        L.ref_medium_from_props.restype = C.c_void_p
        L.ref_medium_from_props.argtypes = [C.c_char_p] * 5 + [C.c_uint64]
        L.ref_kd_count.argtypes = [C.c_void_p, fp, C.c_size_t, C.POINTER(C.c_uint64)]
        L.ref_scene_create_without_sensor.restype = C.c_void_p
        L.ref_scene_create_without_sensor.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(C.c_int)]
        L.ref_sensor_rays_from_props.argtypes = [C.c_char_p, C.c_int, C.c_int, fp, C.c_size_t, fp]
        L.ref_shape_from_props.argtypes = [C.c_char_p, C.c_char_p, fp, C.c_size_t, fp]
        L.ref_defaults.argtypes = [fp]
        L.ref_bsdf_from_props.argtypes = [C.c_char_p, C.c_char_p, C.c_int, fp, fp, fp, C.c_size_t, fp, fp, fp, fp, fp, u32p]
        L.ref_mesh_load.restype = C.c_void_p
        L.ref_mesh_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, fp]
        L.ref_mesh_destroy.argtypes = [C.c_void_p]
        L.ref_mesh_count.argtypes = [C.c_void_p]
        L.ref_mesh_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.ref_mesh_get.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, u32p]
    return _LIB


class RefScene:
    """A scene instantiated from reference objects (PluginManager + Properties) out of the same flat description the product
    and the oracle port consume."""

    def __init__(self, builder=None, desc=None, keep=None, rtrans_reduce=None, without_sensor=False):
        self.L = lib()
        if desc is None:
            # the reduced rough-transmittance tables of the description are not read here: the reference's roughplastic loads
            # data/microfacet/*.dat itself (rtrans.h), so zeros do
            desc, keep = builder.desc(rtrans_reduce=rtrans_reduce or (lambda distr, eta, alpha: (np.zeros(100), 0.0, 0.0)))
        self.desc, self._keep = desc, keep
        if without_sensor:  # the camera is Scene::configure's fallback, the film its default
            wh = (C.c_int * 2)()
            self.h = self.L.ref_scene_create_without_sensor(C.byref(desc), wh)
        else:
            self.h = self.L.ref_scene_create(C.byref(desc))
        if not self.h:
            raise RuntimeError("ref_scene_create failed: %s" % self.L.ref_last_error().decode())
        self.W, self.H = (wh[0], wh[1]) if without_sensor else (desc.film.width, desc.film.height)

    def __del__(self):
        try:
            if self.h:
                self.L.ref_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _ok(self, r):
        if r != 0:
            raise RuntimeError("ref_harness: %s" % self.L.ref_last_error().decode())

    def kd_info(self):
        out = (C.c_uint64 * 2)()
        self._ok(self.L.ref_kd_info(self.h, out))
        return dict(shapes=out[0], prims=out[1])

    def intersect(self, rays, nthreads=0):
        """Scene::rayIntersect(ray, its): dict of t, p, uv, geo_n, sh_n, sh_s, dpdu, prim (global id, 0xFFFFFFFF = miss)."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        out = np.zeros((n, 18), np.float32)
        prim = np.zeros(n, np.uint32)
        self._ok(self.L.ref_intersect(self.h, _f(rays), n, 0, _f(out), _u(prim), nthreads))
        return dict(t=out[:, 0], p=out[:, 1:4], uv=out[:, 4:6], geo_n=out[:, 6:9], sh_n=out[:, 9:12], sh_s=out[:, 12:15],
                    dpdu=out[:, 15:18], prim=prim)

    def occluded(self, rays, nthreads=0):
        """Scene::rayIntersect(ray) (shadow rays): bool per ray."""
        rays = np.ascontiguousarray(rays, np.float32)
        prim = np.zeros(rays.shape[0], np.uint32)
        self._ok(self.L.ref_intersect(self.h, _f(rays), rays.shape[0], 1, None, _u(prim), nthreads))
        return prim == 0

    def kd_count(self, rays):
        """The reference's counting traversal (rayIntersectHavranCollectStatistics) summed over the rays: inner nodes
        traversed, leaf index entries visited, rays that hit."""
        rays = np.ascontiguousarray(rays, np.float32)
        c = (C.c_uint64 * 3)()
        self._ok(self.L.ref_kd_count(self.h, _f(rays), rays.shape[0], c))
        return dict(inner=c[0], indices=c[1], hits=c[2])

    def camera_rays(self, pos):
        pos = np.ascontiguousarray(pos, np.float32)
        rays = np.zeros((pos.shape[0], 8), np.float32)
        self._ok(self.L.ref_camera_rays(self.h, _f(pos), pos.shape[0], _f(rays)))
        return rays

    def bsdf(self, index, wi, wo, u):
        wi, wo, u = (np.ascontiguousarray(a, np.float32) for a in (wi, wo, u))
        n = wi.shape[0]
        ev, pdf, swo = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        w, spdf, fl = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
        self._ok(self.L.ref_bsdf(self.h, index, _f(wi), _f(wo), _f(u), n, _f(ev), _f(pdf), _f(swo), _f(w), _f(spdf), _u(fl)))
        return dict(eval=ev, pdf=pdf, wo=swo, weight=w, spdf=spdf, flags=fl)

    def emitter_sample(self, ref, ref_n, u):
        ref, ref_n, u = (np.ascontiguousarray(a, np.float32) for a in (ref, ref_n, u))
        n = u.shape[0]
        d, dist, pdf, val = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        self._ok(self.L.ref_emitter_direct(self.h, _f(ref), _f(ref_n), _f(u), _f(d), n, _f(dist), _f(pdf), _f(val)))
        return d, dist, pdf, val

    def emitter_pdf(self, ref, ref_n, d):
        ref, ref_n = (np.ascontiguousarray(a, np.float32) for a in (ref, ref_n))
        d = np.ascontiguousarray(d, np.float32).copy()
        pdf = np.zeros(d.shape[0], np.float32)
        self._ok(self.L.ref_emitter_direct(self.h, _f(ref), _f(ref_n), None, _f(d), d.shape[0], None, _f(pdf), None))
        return pdf

    def radiance(self, params, pixel, sample):
        """The integrator's Li for camera samples (pixel, sample index) drawn from the shared counter stream."""
        pixel = np.ascontiguousarray(pixel, np.uint32)
        sample = np.ascontiguousarray(sample, np.uint32)
        out = np.zeros((pixel.shape[0], 3), np.float32)
        pos = np.zeros((pixel.shape[0], 2), np.float32)
        self._ok(self.L.ref_radiance(self.h, C.byref(params), _u(pixel), _u(sample), pixel.shape[0], _f(out), _f(pos)))
        return out, pos

    def film_splat(self, pos, rgb):
        pos, rgb = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(rgb, np.float32)
        film = np.zeros((self.H, self.W, 5), np.float32)
        self._ok(self.L.ref_film_splat(self.h, _f(pos), _f(rgb), pos.shape[0], _f(film)))
        return film

    def render(self, params, first_sample=0, n_samples=1, nthreads=0, independent=False, want_film=True, repeat=1):
        """Scene::preprocess + Scene::render of the reference (every stage but Film::develop). Returns (film H*W*5, seconds of
        the `repeat` Scene::render calls; the film is the last one's)."""
        film = np.zeros((self.H, self.W, 5), np.float32) if want_film else None
        sec, spp = C.c_double(), C.c_int()
        self._ok(self.L.ref_render(self.h, C.byref(params), first_sample, n_samples, _f(film) if want_film else None, nthreads,
                                   int(independent), C.byref(sec), C.byref(spp), repeat))
        self.spp_done = spp.value
        return film, sec.value

    def render_plugin(self, params, plugin, xml_path, device_count=1):
        """The reference's Scene::preprocess / render / postprocess with the integrator its PluginManager loads from
        plugins/<plugin>.so (integration/b200guidedpath.cpp: the reference-side binding of libb200pg.so). Returns the film the
        binding put into the reference's HDRFilm (H*W*5) and the seconds Scene::render took."""
        film = np.zeros((self.H, self.W, 5), np.float32)
        sec = C.c_double()
        self._ok(self.L.ref_render_plugin(self.h, C.byref(params), plugin.encode(), (xml_path or "").encode(), device_count, _f(film),
                                          C.byref(sec)))
        return film, sec.value

    def plugin_put_film(self, rgbaw):
        """integration/b200guidedpath.cpp's film hand-off alone: rgbaw (H, W, 5) -> the reference's HDRFilm -> its storage."""
        rgbaw = np.ascontiguousarray(rgbaw, np.float32)
        film = np.zeros((self.H, self.W, 5), np.float32)
        self._ok(self.L.ref_plugin_put_film(self.h, _f(rgbaw), _f(film)))
        return film

    def grid_lookup(self, medium, p):
        p = np.ascontiguousarray(p, np.float32)
        out = np.zeros(p.shape[0], np.float32)
        self._ok(self.L.ref_grid_lookup(self.h, medium, _f(p), p.shape[0], _f(out)))
        return out

    def medium_sample(self, medium, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        t, ps, tr = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        self._ok(self.L.ref_medium_sample(self.h, medium, _f(rays), n, _f(t), _f(ps), _f(tr)))
        return t, ps, tr

    def phase(self, medium, wi, wo, u):
        wi, wo, u = (np.ascontiguousarray(a, np.float32) for a in (wi, wo, u))
        n = wi.shape[0]
        ev, swo, pdf = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        self._ok(self.L.ref_phase(self.h, medium, _f(wi), _f(wo), _f(u), n, _f(ev), _f(swo), _f(pdf)))
        return ev, swo, pdf


def load_meshes(kind, path, shape_index=0, face_normals=False, flip_normals=False, to_world=None):
    """The reference's own mesh loaders + TriMesh::configure: kind = "obj" (src/shapes/obj.cpp, every sub-mesh) or "serialized"
    (TriMesh(Stream *, shapeIndex), trimesh.cpp:79-270). Returns a list of dicts P (n, 3), N (n, 3) or None, UV (n, 2) or None,
    T (m, 3)."""
    L = lib()
    m = None if to_world is None else np.ascontiguousarray(to_world, np.float32).ravel()
    h = L.ref_mesh_load(kind.encode(), str(path).encode(), shape_index, int(face_normals), int(flip_normals), _f(m) if m is not None else None)
    if not h:
        raise RuntimeError("ref_mesh_load: %s" % L.ref_last_error().decode())
    out = []
    try:
        for i in range(L.ref_mesh_count(h)):
            info = (C.c_uint64 * 4)()
            L.ref_mesh_info(h, i, info)
            nv, nt = int(info[0]), int(info[1])
            P, N, UV, T = np.zeros((nv, 3), np.float32), np.zeros((nv, 3), np.float32), np.zeros((nv, 2), np.float32), np.zeros((nt, 3), np.uint32)
            L.ref_mesh_get(h, i, _f(P), _f(N), _f(UV), _u(T))
            out.append(dict(P=P, N=N if info[2] else None, UV=UV if info[3] else None, T=T))
    finally:
        L.ref_mesh_destroy(h)
    return out


def bsdf_from_props(plugin, props, wi, wo, u, twosided=False):
    """A reference BSDF created from a plugin name and named properties only -- [(name, kind, value)] with kind f / i / b / s / c
    as in an XML scene -- and evaluated like RefScene.bsdf: the reference's own defaults and property semantics."""
    L = lib()
    wi, wo, u = (np.ascontiguousarray(a, np.float32) for a in (wi, wo, u))
    n = wi.shape[0]
    ev, pdf, swo = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    w, spdf, fl = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
    text = ";".join("%s|%s|%s" % t for t in props)
    if L.ref_bsdf_from_props(plugin.encode(), text.encode(), int(twosided), _f(wi), _f(wo), _f(u), n, _f(ev), _f(pdf), _f(swo), _f(w),
                             _f(spdf), _u(fl)) != 0:
        raise RuntimeError("ref_bsdf_from_props: %s" % L.ref_last_error().decode())
    return dict(eval=ev, pdf=pdf, wo=swo, weight=w, spdf=spdf, flags=fl)


def _props_text(props):
    return ";".join("%s|%s|%s" % t for t in props).encode()


def sensor_rays_from_props(props, width, height, pos):
    """A reference `perspective` sensor from named properties only (kind x = transform ops, see ref_harness.cpp parseProps)."""
    L = lib()
    pos = np.ascontiguousarray(pos, np.float32)
    rays = np.zeros((pos.shape[0], 8), np.float32)
    if L.ref_sensor_rays_from_props(_props_text(props), width, height, _f(pos), pos.shape[0], _f(rays)) != 0:
        raise RuntimeError("ref_sensor_rays_from_props: %s" % L.ref_last_error().decode())
    return rays


def shape_hits_from_props(plugin, props, rays):
    """One reference shape plugin from named properties in a ShapeKDTree of its own: hit records like RefScene.intersect."""
    L = lib()
    rays = np.ascontiguousarray(rays, np.float32)
    out = np.zeros((rays.shape[0], 18), np.float32)
    if L.ref_shape_from_props(plugin.encode(), _props_text(props), _f(rays), rays.shape[0], _f(out)) != 0:
        raise RuntimeError("ref_shape_from_props: %s" % L.ref_last_error().decode())
    return dict(t=out[:, 0], p=out[:, 1:4], uv=out[:, 4:6], geo_n=out[:, 6:9], sh_n=out[:, 9:12], sh_s=out[:, 12:15], dpdu=out[:, 15:18])


def defaults():
    out = np.zeros(8, np.float32)
    if lib().ref_defaults(_f(out)) != 0:
        raise RuntimeError("ref_defaults: %s" % lib().ref_last_error().decode())
    return dict(film_width=int(out[0]), film_height=int(out[1]), filter_radius=float(out[2]), sample_count=int(out[3]),
                sampling_weight=float(out[4]))


def medium_from_props(medium, density, albedo, phase_plugin, phase, seed):
    """A reference `heterogeneous` medium assembled from plugin names + named properties (the density gridvolume reads the .vol
    file itself). Returns a RefScene-like handle with grid_lookup / medium_sample / phase on medium 0."""
    L = lib()
    self = RefScene.__new__(RefScene)
    self.L, self.desc, self._keep = L, None, None
    self.h = L.ref_medium_from_props(_props_text(medium), _props_text(density), _props_text(albedo), (phase_plugin or "").encode(),
                                     _props_text(phase), seed)
    if not self.h:
        raise RuntimeError("ref_medium_from_props: %s" % L.ref_last_error().decode())
    self.W = self.H = 0
    return self
