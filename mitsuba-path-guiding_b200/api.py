"""Python host mirror of the C-ABI (include/b200pg.h) via ctypes.

Mirrors the reference's object model for this path: ``Scene`` (scene XML or flat arrays) and
``Integrator`` (``progressivepath`` / ``progressivevolpath`` parameters, ``render()``,
``cancel()``, film access). No CPU fallback: a missing library or device raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200PG_LIB: development aid (tools/build_variant.sh builds the library with other compile-time knobs for A/B runs)
LIB_PATH = os.environ.get("B200PG_LIB") or os.path.join(_HERE, "libb200pg.so")

fp = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)

_lib = None


class B200pgError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200pgError("libb200pg.so not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    L.b200pg_version.restype = C.c_int
    L.b200pg_last_error.restype = C.c_char_p
    L.b200pg_integrator_params_default.argtypes = [C.POINTER(A.IntegratorParams)]
    L.b200pg_scene_load_xml.restype = C.c_void_p
    L.b200pg_scene_load_xml.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_char_p, C.c_size_t]
    L.b200pg_scene_from_arrays.restype = C.c_void_p
    L.b200pg_scene_from_arrays.argtypes = [C.POINTER(A.SceneDesc)]
    L.b200pg_scene_desc.restype = C.POINTER(A.SceneDesc)
    L.b200pg_scene_desc.argtypes = [C.c_void_p]
    L.b200pg_scene_integrator_params.argtypes = [C.c_void_p, C.POINTER(A.IntegratorParams)]
    L.b200pg_scene_destroy.argtypes = [C.c_void_p]
    L.b200pg_integrator_create.restype = C.c_void_p
    L.b200pg_integrator_create.argtypes = [C.c_void_p, C.POINTER(A.IntegratorParams), C.c_int]
    L.b200pg_render.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    for name in ("b200pg_cancel", "b200pg_film_clear", "b200pg_train_accumulate", "b200pg_train_end"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.b200pg_train_update.argtypes = [C.c_void_p, C.c_int]
    L.b200pg_train_begin.argtypes = [C.c_void_p, u32p, u32p]
    L.b200pg_guiding_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.b200pg_progression_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.b200pg_train_stats_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.b200pg_film_device_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.b200pg_film_read.argtypes = [C.c_void_p, fp]
    L.b200pg_film_develop.argtypes = [C.c_void_p, fp]
    L.b200pg_film_write.argtypes = [C.c_void_p, C.c_char_p]
    L.b200pg_stats.argtypes = [C.c_void_p, C.POINTER(A.Stats)]
    L.b200pg_destroy.argtypes = [C.c_void_p]
    L.b200pg_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.b200pg_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.b200pg_scene_upload.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    L.b200pg_k_trace.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_int, fp, u32p]
    L.b200pg_k_trace_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, fp, C.POINTER(C.c_uint64)]
    L.b200pg_k_bsdf.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_size_t, fp, fp, fp, fp, fp, u32p]
    L.b200pg_k_radiance.argtypes = [C.c_void_p, u32p, u32p, C.c_size_t, fp]
    L.b200pg_k_film_splat.argtypes = [C.c_void_p, fp, fp, C.c_size_t]
    L.b200pg_k_grid_lookup.argtypes = [C.c_void_p, C.c_int, fp, C.c_size_t, fp]
    L.b200pg_k_medium_sample.argtypes = [C.c_void_p, C.c_int, fp, C.c_size_t, fp, fp, fp, fp]
    L.b200pg_k_vmm_pdf_sample.argtypes = [C.c_void_p, fp, fp, fp, C.c_size_t, fp, fp, fp, u32p]
    L.b200pg_k_bin_samples.argtypes = [C.c_void_p, fp, C.c_size_t, u32p, u32p, u32p, u32p]
    L.b200pg_k_em_step.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, C.c_size_t, C.c_int, fp]
    L.b200pg_film_read_async.argtypes = [C.c_void_p, fp]
    L.b200pg_film_read_wait.argtypes = [C.c_void_p]
    L.b200pg_train.argtypes = [C.c_void_p, C.c_int, u32p, u32p]
    L.b200pg_comm_local_handle.argtypes = [C.c_void_p, C.c_void_p]
    L.b200pg_comm_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.b200pg_field_snapshot.argtypes = [C.c_void_p, u32p, C.POINTER(C.c_size_t)]
    L.b200pg_field_load.argtypes = [C.c_void_p, u32p, C.c_size_t]
    L.b200pg_features_read.argtypes = [C.c_void_p, fp]
    L.b200pg_features_write.argtypes = [C.c_void_p, C.c_char_p]
    L.b200pg_film_ipc_handle.argtypes = [C.c_void_p, C.c_void_p]
    L.b200pg_film_add_peers.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.b200pg_film_peers_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.b200pg_k_em_exchange.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_float)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise B200pgError((lib().b200pg_last_error() or b"").decode() or "b200pg error %d" % rc)


def _f(a):
    return a.ctypes.data_as(fp)


def _u(a):
    return a.ctypes.data_as(u32p)


def default_params():
    p = A.IntegratorParams()
    lib().b200pg_integrator_params_default(C.byref(p))
    return p


class Scene:
    """Scene handle: ``Scene.from_builder(SceneBuilder)`` or ``Scene.load_xml(path, defines)``."""

    def __init__(self, handle, keep=None):
        self.h = handle
        self._keep = keep

    @classmethod
    def from_builder(cls, builder):
        def no_tables(distr, eta, alpha):  # let the C++ loader reduce the tables itself
            return np.zeros(100, np.float32), 0.0, 0.0

        desc, keep = builder.desc(rtrans_reduce=no_tables)
        h = lib().b200pg_scene_from_arrays(C.byref(desc))
        if not h:
            raise B200pgError(lib().b200pg_last_error().decode())
        return cls(h, keep)

    @classmethod
    def load_xml(cls, path, defines=None):
        arr = None
        if defines:
            items = [("%s=%s" % kv).encode() for kv in defines.items()]
            arr = (C.c_char_p * (len(items) + 1))(*items, None)
        err = C.create_string_buffer(1024)
        h = lib().b200pg_scene_load_xml(path.encode(), arr, err, 1024)
        if not h:
            raise B200pgError(err.value.decode() or lib().b200pg_last_error().decode())
        return cls(h)

    @property
    def desc(self):
        return lib().b200pg_scene_desc(self.h).contents

    def integrator_params(self):
        p = A.IntegratorParams()
        _check(lib().b200pg_scene_integrator_params(self.h, C.byref(p)))
        return p

    def close(self):
        if self.h:
            lib().b200pg_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Integrator:
    def __init__(self, scene, params=None, device=0):
        self.scene = scene
        self.params = params if params is not None else scene.integrator_params()
        self.h = lib().b200pg_integrator_create(scene.h, C.byref(self.params), device)
        if not self.h:
            raise B200pgError(lib().b200pg_last_error().decode())
        d = scene.desc
        self.W, self.H = d.film.width, d.film.height

    def close(self):
        if self.h:
            lib().b200pg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- rendering
    def render(self, devices=None):
        """Integrator::render. devices = list of CUDA device indices (first = this integrator's own) renders on all of them
        inside the one call; None = this integrator's device."""
        if devices:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            _check(lib().b200pg_render(self.h, len(devices), arr))
        else:
            _check(lib().b200pg_render(self.h, 0, None))

    def cancel(self):
        _check(lib().b200pg_cancel(self.h))

    def progression(self, first_sample, n_samples, rows=None):
        r0, r1 = rows if rows else (0, 0)
        _check(lib().b200pg_progression_render(self.h, first_sample, n_samples, r0, r1))

    def film_clear(self):
        _check(lib().b200pg_film_clear(self.h))

    def film(self, out=None):
        """Raw accumulators (H, W, 5) = R, G, B, alpha, weight. Pass a pinned ``out`` array for a fast copy."""
        if out is None:
            out = np.empty((self.H, self.W, 5), np.float32)
        assert out.dtype == np.float32 and out.size == self.H * self.W * 5 and out.flags["C_CONTIGUOUS"]
        _check(lib().b200pg_film_read(self.h, _f(out)))
        return out

    def film_async(self, out):
        """Snapshot the film into the (pinned) array `out` while rendering continues; film_wait() blocks until it is there."""
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        _check(lib().b200pg_film_read_async(self.h, _f(out)))

    def film_wait(self):
        _check(lib().b200pg_film_read_wait(self.h))

    def film_ipc_handle(self):
        """64-byte CUDA IPC handle of this integrator's film (for the other ranks' film_peers_connect / film_add_peers)."""
        buf = C.create_string_buffer(64)
        _check(lib().b200pg_film_ipc_handle(self.h, buf))
        return buf.raw

    def film_add_peers(self, rank, world, handles):
        """film += the other ranks' films, read over NVLink (end of a multi-GPU job)."""
        buf = C.create_string_buffer(bytes(handles), 64 * world)
        _check(lib().b200pg_film_add_peers(self.h, rank, world, buf))

    def film_peers_connect(self, rank, world, handles):
        """Map the other ranks' films (world x 64-byte IPC handles): film_async then delivers the job's merged film."""
        buf = C.create_string_buffer(bytes(handles), 64 * world) if handles else None
        _check(lib().b200pg_film_peers_connect(self.h, rank, world, buf))

    def develop(self):
        out = np.zeros((self.H, self.W, 3), np.float32)
        _check(lib().b200pg_film_develop(self.h, _f(out)))
        return out

    def features(self):
        """Denoiser feature buffers (H, W, 10) = color.rgb, albedo.rgb, normal.xyz, sample count (set_option("feature_buffers", 1))."""
        out = np.zeros((self.H, self.W, 10), np.float32)
        _check(lib().b200pg_features_read(self.h, _f(out)))
        return out

    def features_write(self, path):
        _check(lib().b200pg_features_write(self.h, path.encode()))

    def film_write(self, path):
        _check(lib().b200pg_film_write(self.h, path.encode()))

    def film_device_buffer(self):
        p = C.c_void_p()
        n = C.c_size_t()
        _check(lib().b200pg_film_device_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stats(self):
        s = A.Stats()
        _check(lib().b200pg_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in A.Stats._fields_}

    def set_option(self, name, value):
        _check(lib().b200pg_set_option(self.h, name.encode(), int(value)))

    def stage_times(self):
        sec = (C.c_double * 5)()
        cnt = (C.c_uint64 * 5)()
        _check(lib().b200pg_stage_times(self.h, sec, cnt))
        names = ("trace", "shade", "shadow", "film", "train")
        return {n: dict(seconds=sec[i], launches=cnt[i]) for i, n in enumerate(names)}

    def scene_upload(self):
        n = C.c_size_t()
        _check(lib().b200pg_scene_upload(self.h, C.byref(n)))
        return n.value

    # ---- guiding / training hooks
    def guiding_mode(self, record, sample):
        _check(lib().b200pg_guiding_mode(self.h, int(record), int(sample)))

    def train_begin(self):
        n, c = C.c_uint32(), C.c_uint32()
        _check(lib().b200pg_train_begin(self.h, C.byref(n), C.byref(c)))
        return n.value, c.value

    def train_accumulate(self):
        _check(lib().b200pg_train_accumulate(self.h))

    def train_stats_buffer(self):
        p = C.c_void_p()
        n = C.c_size_t()
        _check(lib().b200pg_train_stats_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def train_update(self, commit):
        _check(lib().b200pg_train_update(self.h, int(commit)))

    def train_end(self):
        _check(lib().b200pg_train_end(self.h))

    def train(self, n_iter=4, allreduce=None):
        """One training update; ``allreduce(dev_ptr, n_floats)`` sums the statistics buffer over ranks (optional)."""
        n, c = self.train_begin()
        for it in range(n_iter):
            self.train_accumulate()
            if allreduce is not None:
                allreduce(*self.train_stats_buffer())
            self.train_update(it == n_iter - 1)
        self.train_end()
        return n, c

    def train_fused(self, n_iter=4):
        """One training update in a single library call; sums statistics over the connected ranks (comm_connect)."""
        n, c = C.c_uint32(0), C.c_uint32(0)
        _check(lib().b200pg_train(self.h, n_iter, C.byref(n), C.byref(c)))
        return n.value, c.value

    def comm_local_handle(self):
        buf = (C.c_ubyte * 64)()
        _check(lib().b200pg_comm_local_handle(self.h, buf))
        return bytes(buf)

    def comm_connect(self, rank, world, handles):
        assert len(handles) == 64 * world
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        _check(lib().b200pg_comm_connect(self.h, rank, world, buf))

    def em_exchange_bench(self, n_cells, n_iter=20, mode=0):
        """Average ms per fused statistics sum + M-step launch over n_cells synthetic cells (b200pg_k_em_exchange);
        mode 0 = as trained, 1 = local M-step only, 2 = all-read form, 3 = reduce-scatter + all-gather form."""
        ms = C.c_float(0)
        _check(lib().b200pg_k_em_exchange(self.h, n_cells, n_iter, int(mode), C.byref(ms)))
        return ms.value

    def field_snapshot(self):
        n = C.c_size_t(0)
        _check(lib().b200pg_field_snapshot(self.h, None, C.byref(n)))
        w = np.zeros(n.value, np.uint32)
        _check(lib().b200pg_field_snapshot(self.h, _u(w), C.byref(n)))
        return w

    def field_load(self, words):
        words = np.ascontiguousarray(words, np.uint32)
        _check(lib().b200pg_field_load(self.h, _u(words), words.size))

    def k_vmm_pdf_sample(self, pos, dirs, u):
        pos, dirs, u = (np.ascontiguousarray(x, np.float32) for x in (pos, dirs, u))
        n = pos.shape[0]
        pdf, sd, spdf, cell = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
        _check(lib().b200pg_k_vmm_pdf_sample(self.h, _f(pos), _f(dirs), _f(u), n, _f(pdf), _f(sd), _f(spdf), _u(cell)))
        return dict(pdf=pdf, dir=sd, spdf=spdf, cell=cell)

    def k_bin_samples(self, pos, n_cells):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        cell, perm, off = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n_cells + 1, np.uint32)
        nc = C.c_uint32()
        _check(lib().b200pg_k_bin_samples(self.h, _f(pos), n, _u(cell), _u(perm), _u(off), C.byref(nc)))
        assert nc.value == n_cells
        return cell, perm, off

    def k_em_step(self, s, n_iter, n_cells, K):
        a = [np.ascontiguousarray(s[k], np.float32) for k in ("pos", "dir", "weight", "pdf", "dist")]
        st = np.zeros(n_cells * (4 * K + 8), np.float32)
        _check(lib().b200pg_k_em_step(self.h, *[_f(x) for x in a], a[0].shape[0], n_iter, _f(st)))
        return st.reshape(n_cells, 4 * K + 8)

    # ---- per-kernel entry points
    def k_trace(self, rays, shadow=False):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        tuv = np.zeros((n, 3), np.float32)
        prim = np.zeros(n, np.uint32)
        _check(lib().b200pg_k_trace(self.h, _f(rays), n, int(shadow), _f(tuv), _u(prim)))
        return tuv, prim

    def k_trace_device(self, d_rays, n, d_hits, shadow=False, count=False):
        ms = C.c_float()
        cnt = (C.c_uint64 * 2)()
        _check(lib().b200pg_k_trace_device(self.h, d_rays, n, int(shadow), d_hits, C.byref(ms), cnt if count else None))
        return ms.value, (cnt[0], cnt[1])

    def k_bsdf(self, index, wi, wo, u):
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
        _check(lib().b200pg_k_bsdf(self.h, index, _f(wi), _f(wo), _f(u), n, _f(ev), _f(pdf), _f(swo), _f(w), _f(spdf), _u(fl)))
        return dict(eval=ev, pdf=pdf, wo=swo, weight=w, spdf=spdf, flags=fl)

    def k_radiance(self, pixel, sample):
        pixel = np.ascontiguousarray(pixel, np.uint32)
        sample = np.ascontiguousarray(sample, np.uint32)
        out = np.zeros((pixel.shape[0], 3), np.float32)
        _check(lib().b200pg_k_radiance(self.h, _u(pixel), _u(sample), pixel.shape[0], _f(out)))
        return out

    def k_film_splat(self, pos, rgb):
        pos = np.ascontiguousarray(pos, np.float32)
        rgb = np.ascontiguousarray(rgb, np.float32)
        _check(lib().b200pg_k_film_splat(self.h, _f(pos), _f(rgb), pos.shape[0]))

    def k_grid_lookup(self, medium, p):
        p = np.ascontiguousarray(p, np.float32)
        out = np.zeros(p.shape[0], np.float32)
        _check(lib().b200pg_k_grid_lookup(self.h, medium, _f(p), p.shape[0], _f(out)))
        return out

    def k_medium_sample(self, medium, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        t, tr, wo, pdf = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        _check(lib().b200pg_k_medium_sample(self.h, medium, _f(rays), n, _f(t), _f(tr), _f(wo), _f(pdf)))
        return t, tr, wo, pdf
