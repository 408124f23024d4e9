"""ctypes wrapper around oracle/_build/liboracle.so (TEST INFRASTRUCTURE: the CPU oracle)."""
import ctypes as C
import os
import sys

import numpy as np

from conftest import ROOT, load_package

b200pg = load_package()
A = b200pg._abi

fp = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def _f(a):
    return a.ctypes.data_as(fp)


def _u(a):
    return a.ctypes.data_as(u32p)


class Oracle:
    def __init__(self, so=None):
        so = so or os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        self.lib = L = C.CDLL(so)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.POINTER(A.SceneDesc)]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_kd_info.argtypes = [C.c_void_p, u64p]
        L.orc_trace.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_int, fp, u32p, u64p, C.c_int]
        L.orc_trace_bruteforce.argtypes = [C.c_void_p, fp, C.c_size_t, fp, u32p]
        L.orc_camera_rays.argtypes = [C.c_void_p, fp, C.c_size_t, fp]
        L.orc_bsdf.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_size_t, fp, fp, fp, fp, fp, u32p]
        L.orc_radiance.argtypes = [C.c_void_p, C.POINTER(A.IntegratorParams), u32p, u32p, C.c_size_t, fp, C.c_void_p, C.c_void_p]
        L.orc_film_splat.argtypes = [C.c_void_p, fp, fp, C.c_size_t, fp]
        L.orc_clipped_aabb.argtypes = [fp, fp, fp]
        L.orc_grid_lookup.argtypes = [C.c_void_p, C.c_int, fp, C.c_size_t, fp]
        L.orc_medium_sample.argtypes = [C.c_void_p, C.c_int, fp, C.c_size_t, fp, fp, fp, fp]
        L.orc_phase.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_size_t, fp, fp, fp]
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(A.IntegratorParams), C.c_int, C.c_int, C.c_int, C.c_int, fp,
                                 C.c_int, u64p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.orc_emitter_direct.argtypes = [C.c_void_p, fp, fp, fp, fp, C.c_size_t, fp, fp, fp]
        L.orc_intersect.argtypes = [C.c_void_p, fp, C.c_size_t, fp]
        L.orc_microfacet.argtypes = [C.c_int, C.c_float, C.c_float, fp, fp, fp, C.c_size_t, fp, fp, fp]
        L.orc_features.argtypes = [C.c_void_p, C.POINTER(A.IntegratorParams), C.c_int, C.c_int, fp]
        L.orc_field_create.restype = C.c_void_p
        L.orc_field_create.argtypes = [C.c_int, fp, fp]
        L.orc_field_destroy.argtypes = [C.c_void_p]
        L.orc_field_snapshot.restype = C.c_size_t
        L.orc_field_snapshot.argtypes = [C.c_void_p, u32p, C.c_size_t]
        L.orc_field_load.argtypes = [C.c_void_p, u32p, C.c_size_t]
        L.orc_field_info.argtypes = [C.c_void_p, u32p]
        L.orc_vmm_pdf_sample.argtypes = [C.c_void_p, fp, fp, fp, C.c_size_t, fp, fp, fp, u32p]
        L.orc_bin_samples.argtypes = [C.c_void_p, fp, C.c_size_t, u32p, u32p, u32p]
        L.orc_estep.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, C.c_size_t, fp]
        L.orc_train.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, C.c_size_t, C.c_int, C.c_float]
        L.orc_train_sink.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        L.orc_train_levels.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, C.c_size_t, C.c_int, C.c_float, C.c_int]
        L.orc_samples_create.restype = C.c_void_p
        L.orc_samples_destroy.argtypes = [C.c_void_p]
        L.orc_samples_size.restype = C.c_size_t
        L.orc_samples_size.argtypes = [C.c_void_p]
        L.orc_samples_clear.argtypes = [C.c_void_p]
        L.orc_samples_get.argtypes = [C.c_void_p, fp, fp, fp, fp, fp]
        L.orc_rtrans_reduce.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_float, fp, fp, fp]
        L.orc_num_threads.restype = C.c_int

    def num_threads(self):
        return self.lib.orc_num_threads()

    # ---- rough transmittance tables
    def microfacet(self, distr, alpha_u, alpha_v, wi, u=None, m=None):
        """MicrofacetDistribution (microfacet.h): visible-normal samples for u (n, 2), or the given normals m (n, 3);
        returns m, pdfVisible(wi, m), D(m), smithG1(wi, m)."""
        wi = np.ascontiguousarray(wi, np.float32)
        if u is not None:
            u = np.ascontiguousarray(u, np.float32)
            n = u.shape[0]
            m = np.zeros((n, 3), np.float32)
        else:
            m = np.ascontiguousarray(m, np.float32).copy()
            n = m.shape[0]
        pdf, D, G1 = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.lib.orc_microfacet(int(distr), alpha_u, alpha_v, _f(wi), _f(u) if u is not None else None, _f(m), n, _f(pdf), _f(D), _f(G1))
        return m, pdf, D, G1

    def clipped_aabb(self, tri, box_min, box_max):
        t = np.ascontiguousarray(tri, np.float32).ravel()
        b = np.ascontiguousarray(list(box_min) + list(box_max), np.float32)
        out = np.zeros(6, np.float32)
        self.lib.orc_clipped_aabb(_f(t), _f(b), _f(out))
        return out[:3], out[3:]

    def rtrans_reduce(self, table, eta, alpha):
        """table: dict from b200pg.rtrans.load_packed / load_dat (raw floats in file order)."""
        ext = np.zeros(table["thetaN"], np.float32)
        ed = C.c_float()
        idf = C.c_float()
        raw = np.ascontiguousarray(table["raw"], np.float32)
        self.lib.orc_rtrans_reduce(_f(raw), table["etaN"], table["alphaN"], table["thetaN"], table["etaMin"],
                                   table["etaMax"], table["alphaMin"], table["alphaMax"], eta, alpha, _f(ext),
                                   C.byref(ed), C.byref(idf))
        return ext, ed.value, idf.value

    def scene(self, builder):
        return OracleScene(self, builder)

    def field(self, K, bmin, bmax):
        return OracleField(self, K, bmin, bmax)

    def samples(self):
        return OracleSamples(self)


class OracleSamples:
    """Training-sample sink filled by guided/unguided oracle renders."""

    def __init__(self, orc):
        self.L = orc.lib
        self.h = self.L.orc_samples_create()

    def __del__(self):
        try:
            self.L.orc_samples_destroy(self.h)
        except Exception:
            pass

    def clear(self):
        self.L.orc_samples_clear(self.h)

    def get(self):
        n = self.L.orc_samples_size(self.h)
        pos, dirs = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        w, pdf, dist = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        if n:
            self.L.orc_samples_get(self.h, _f(pos), _f(dirs), _f(w), _f(pdf), _f(dist))
        return dict(pos=pos, dir=dirs, weight=w, pdf=pdf, dist=dist)


class OracleField:
    """CPU statement of the guiding field (oracle/oracle_guiding.h)."""

    def __init__(self, orc, K, bmin, bmax):
        self.L = orc.lib
        a, b = np.asarray(bmin, np.float32), np.asarray(bmax, np.float32)
        self.h = self.L.orc_field_create(K, _f(a), _f(b))

    def __del__(self):
        try:
            self.L.orc_field_destroy(self.h)
        except Exception:
            pass

    def info(self):
        out = (C.c_uint32 * 3)()
        self.L.orc_field_info(self.h, out)
        return dict(nodes=out[0], cells=out[1], K=out[2])

    def snapshot(self):
        n = self.L.orc_field_snapshot(self.h, None, 0)
        w = np.zeros(n, np.uint32)
        self.L.orc_field_snapshot(self.h, _u(w), n)
        return w

    def load(self, words):
        words = np.ascontiguousarray(words, np.uint32)
        assert self.L.orc_field_load(self.h, _u(words), words.size) == 0

    def pdf_sample(self, pos, dirs, u):
        pos, dirs, u = (np.ascontiguousarray(x, np.float32) for x in (pos, dirs, u))
        n = pos.shape[0]
        pdf, sd, spdf, cell = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
        self.L.orc_vmm_pdf_sample(self.h, _f(pos), _f(dirs), _f(u), n, _f(pdf), _f(sd), _f(spdf), _u(cell))
        return dict(pdf=pdf, dir=sd, spdf=spdf, cell=cell)

    def bin(self, pos):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        cell, perm = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        off = np.zeros(self.info()["cells"] + 1, np.uint32)
        self.L.orc_bin_samples(self.h, _f(pos), n, _u(cell), _u(perm), _u(off))
        return cell, perm, off

    def _args(self, s):
        a = [np.ascontiguousarray(s[k], np.float32) for k in ("pos", "dir", "weight", "pdf", "dist")]
        return a, a[0].shape[0]

    def estep(self, s):
        a, n = self._args(s)
        i = self.info()
        st = np.zeros(i["cells"] * (4 * i["K"] + 8), np.float32)
        self.L.orc_estep(self.h, *[_f(x) for x in a], n, _f(st))
        return st.reshape(i["cells"], 4 * i["K"] + 8)

    def train_sink(self, sink, n_iter=4, max_cell_samples=32768):
        self.L.orc_train_sink(self.h, sink.h, n_iter, max_cell_samples)

    def train(self, s, n_iter=4, max_cell_samples=32768, split_levels=1):
        a, n = self._args(s)
        if split_levels > 1:
            self.L.orc_train_levels(self.h, *[_f(x) for x in a], n, n_iter, max_cell_samples, split_levels)
        else:
            self.L.orc_train(self.h, *[_f(x) for x in a], n, n_iter, max_cell_samples)


class OracleScene:
    def __init__(self, orc, builder):
        self.orc = orc
        self.L = orc.lib
        from b200pg import rtrans

        def red(distr, eta, alpha):
            return orc.rtrans_reduce(rtrans.load_packed(distr), eta, alpha)

        self.desc, self._keep = builder.desc(rtrans_reduce=red)
        self.h = self.L.orc_scene_create(C.byref(self.desc))
        if not self.h:
            raise RuntimeError("orc_scene_create failed")
        self.W, self.H = builder.width, builder.height

    @classmethod
    def from_desc(cls, orc, desc, keep=None):
        """Oracle scene from a flat description (e.g. api.Scene.load_xml(...).desc: what the product's XML reader produced)."""
        self = cls.__new__(cls)
        self.orc, self.L = orc, orc.lib
        self.desc, self._keep = desc, keep
        self.h = self.L.orc_scene_create(C.byref(desc))
        if not self.h:
            raise RuntimeError("orc_scene_create failed")
        self.W, self.H = desc.film.width, desc.film.height
        return self

    def __del__(self):
        try:
            if self.h:
                self.L.orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def kd_info(self):
        out = (C.c_uint64 * 3)()
        self.L.orc_kd_info(self.h, out)
        return dict(nodes=out[0], indices=out[1], prims=out[2])

    def trace(self, rays, shadow=False, nthreads=0):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        tuv = np.zeros((n, 3), np.float32)
        prim = np.zeros(n, np.uint32)
        cnt = (C.c_uint64 * 4)()
        self.L.orc_trace(self.h, _f(rays), n, int(shadow), _f(tuv), _u(prim), cnt, nthreads)
        return tuv, prim, dict(nodes=cnt[0], indices=cnt[1], prims=cnt[2], leaves=cnt[3])

    def intersect(self, rays):
        """Full intersection records: dict of t, p, uv, geo_n, sh_n, sh_s, dpdu (skdtree.h:343-428)."""
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.zeros((rays.shape[0], 18), np.float32)
        self.L.orc_intersect(self.h, _f(rays), rays.shape[0], _f(out))
        return dict(t=out[:, 0], p=out[:, 1:4], uv=out[:, 4:6], geo_n=out[:, 6:9], sh_n=out[:, 9:12], sh_s=out[:, 12:15],
                    dpdu=out[:, 15:18])

    def emitter_sample(self, ref, ref_n, u):
        """Scene::sampleEmitterDirect without the visibility test: directions, distances, solid-angle pdfs, radiance / pdf."""
        ref, ref_n, u = (np.ascontiguousarray(a, np.float32) for a in (ref, ref_n, u))
        n = u.shape[0]
        d, dist, pdf, val = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        self.L.orc_emitter_direct(self.h, _f(ref), _f(ref_n), _f(u), _f(d), n, _f(dist), _f(pdf), _f(val))
        return d, dist, pdf, val

    def emitter_pdf(self, ref, ref_n, d):
        """Scene::pdfEmitterDirect for rays (ref, d) whose first hit is an emitter (0 otherwise)."""
        ref, ref_n = (np.ascontiguousarray(a, np.float32) for a in (ref, ref_n))
        d = np.ascontiguousarray(d, np.float32).copy()
        n = d.shape[0]
        pdf = np.zeros(n, np.float32)
        self.L.orc_emitter_direct(self.h, _f(ref), _f(ref_n), None, _f(d), n, None, _f(pdf), None)
        return pdf

    def trace_bruteforce(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        tuv = np.zeros((n, 3), np.float32)
        prim = np.zeros(n, np.uint32)
        self.L.orc_trace_bruteforce(self.h, _f(rays), n, _f(tuv), _u(prim))
        return tuv, prim

    def camera_rays(self, pos):
        pos = np.ascontiguousarray(pos, np.float32)
        rays = np.zeros((pos.shape[0], 8), np.float32)
        self.L.orc_camera_rays(self.h, _f(pos), pos.shape[0], _f(rays))
        return rays

    def bsdf(self, index, wi, wo, u):
        wi = np.ascontiguousarray(wi, np.float32)
        wo = np.ascontiguousarray(wo, np.float32)
        u = np.ascontiguousarray(u, np.float32)
        n = wi.shape[0]
        ev = np.zeros((n, 3), np.float32)
        pdf = np.zeros(n, np.float32)
        swo = np.zeros((n, 3), np.float32)
        w = np.zeros((n, 3), np.float32)
        spdf = np.zeros(n, np.float32)
        fl = np.zeros(n, np.uint32)
        r = self.L.orc_bsdf(self.h, index, _f(wi), _f(wo), _f(u), n, _f(ev), _f(pdf), _f(swo), _f(w), _f(spdf), _u(fl))
        assert r == 0
        return dict(eval=ev, pdf=pdf, wo=swo, weight=w, spdf=spdf, flags=fl)

    def radiance(self, params, pixel, sample, field=None, sink=None):
        pixel = np.ascontiguousarray(pixel, np.uint32)
        sample = np.ascontiguousarray(sample, np.uint32)
        out = np.zeros((pixel.shape[0], 3), np.float32)
        self.L.orc_radiance(self.h, C.byref(params), _u(pixel), _u(sample), pixel.shape[0], _f(out),
                            field.h if field is not None else None, sink.h if sink is not None else None)
        return out

    def grid_lookup(self, medium, p):
        p = np.ascontiguousarray(p, np.float32)
        out = np.zeros(p.shape[0], np.float32)
        assert self.L.orc_grid_lookup(self.h, medium, _f(p), p.shape[0], _f(out)) == 0
        return out

    def medium_sample(self, medium, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        t, tr, wo, pdf = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        assert self.L.orc_medium_sample(self.h, medium, _f(rays), n, _f(t), _f(tr), _f(wo), _f(pdf)) == 0
        return t, tr, wo, pdf

    def phase(self, medium, wi, wo, u):
        wi, wo, u = (np.ascontiguousarray(a, np.float32) for a in (wi, wo, u))
        n = wi.shape[0]
        ev, swo, pdf = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        assert self.L.orc_phase(self.h, medium, _f(wi), _f(wo), _f(u), n, _f(ev), _f(swo), _f(pdf)) == 0
        return ev, swo, pdf

    def film_splat(self, pos, rgb):
        pos = np.ascontiguousarray(pos, np.float32)
        rgb = np.ascontiguousarray(rgb, np.float32)
        film = np.zeros((self.H, self.W, 5), np.float32)
        self.L.orc_film_splat(self.h, _f(pos), _f(rgb), pos.shape[0], _f(film))
        return film

    def features(self, params, first_sample=0, n_samples=1, out=None):
        """Denoiser feature buffers (denoiser.cpp:138-144): (H, W, 10) running means {color, albedo, normal, count}."""
        if out is None:
            out = np.zeros((self.H, self.W, 10), np.float32)
        assert self.L.orc_features(self.h, C.byref(params), first_sample, n_samples, _f(out)) == 0
        return out

    def render(self, params, first_sample=0, n_samples=1, rows=None, film=None, nthreads=0, field=None, sink=None):
        if film is None:
            film = np.zeros((self.H, self.W, 5), np.float32)
        st = (C.c_uint64 * 7)()
        sec = C.c_double()
        r0, r1 = rows if rows else (0, self.H)
        self.L.orc_render(self.h, C.byref(params), first_sample, n_samples, r0, r1, _f(film), nthreads, st,
                          C.byref(sec), field.h if field is not None else None, sink.h if sink is not None else None)
        stats = dict(paths=st[0], normal_rays=st[1], shadow_rays=st[2], path_length_sum=st[3], kd_nodes=st[4],
                     kd_indices=st[5], prim_tests=st[6], seconds=sec.value)
        return film, stats


def develop(film):
    """RGB / weight (fmtconv.cpp:978-1005)."""
    w = film[..., 4:5]
    return np.where(w > 0, film[..., :3] / np.maximum(w, 1e-30), 0.0).astype(np.float32)
