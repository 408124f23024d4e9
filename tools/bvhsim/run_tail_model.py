"""Host replay behind the visit budget + warp-cooperative kernel (log 15): on the full C4 mesh, how long is the
dependent chain of a ray in the sequential walk (node visits + leaf visits: one memory round trip each) and in the cooperative
schedule of k_trace_tail (rounds of up to 32 entries)? No GPU involved. usage: python tools/bvhsim/run_tail_model.py [n_mesh]
Writes nothing; prints the table kept in profiles/r02_bvhsim_tail.txt."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from b200pg import api  # noqa: E402

so = "/tmp/libbvhsim.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-o", so,
                       os.path.join(HERE, "bvhsim.cpp")])
L = C.CDLL(so)
u32p, f32p = C.POINTER(C.c_uint32), C.POINTER(C.c_float)
L.bvhsim_trace.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_int, C.c_int, u32p, u32p, u32p, f32p, u32p]
L.bvhsim_coop.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_int, C.c_int, u32p, u32p, u32p, f32p, u32p]


def fp(a):
    return a.ctypes.data_as(f32p)


def up(a):
    return a.ctypes.data_as(u32p)


n_mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 2237
sb = pkg.scenes.mesh_scene(256, 256, n=n_mesh)
scene = api.Scene.from_builder(sb)
rng = np.random.RandomState(43)
sets = {}
n = 200000
a = rng.randn(n, 3); a = 1.5 * a / np.linalg.norm(a, axis=1, keepdims=True)
b = rng.randn(n, 3); b = 1.5 * b / np.linalg.norm(b, axis=1, keepdims=True)
d = b - a; ln = np.linalg.norm(d, axis=1, keepdims=True); d /= ln
sets["chords through the bounding sphere (test_kd.cpp pattern)"] = np.concatenate([a, np.zeros((n, 1)), d, ln], 1).astype(np.float32)
phi = rng.rand(n) * 2 * np.pi
o = np.stack([1.6 * np.cos(phi), rng.uniform(-0.04, 0.06, n), 1.6 * np.sin(phi)], 1)
t = np.stack([rng.uniform(-0.9, 0.9, n), rng.uniform(-0.05, 0.05, n), rng.uniform(-0.9, 0.9, n)], 1)
d = t - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
sets["rays skimming the sheet"] = np.concatenate([o, np.zeros((n, 1)), d, np.full((n, 1), 10.0)], 1).astype(np.float32)
print("mesh n = %d (%.1f M triangles); closest-hit queries; budget of the first kernel: 96 node visits" % (n_mesh, 2 * (n_mesh - 1) ** 2 / 1e6))
for name, rays in sets.items():
    m = rays.shape[0]
    ns, pt, lv = (np.zeros(m, np.uint32) for _ in range(3))
    t0, p0 = np.zeros(m, np.float32), np.zeros(m, np.uint32)
    L.bvhsim_trace(scene.h, fp(rays), m, 0, 0, up(ns), up(pt), up(lv), fp(t0), up(p0))
    rd, nt, ms = (np.zeros(m, np.uint32) for _ in range(3))
    t1, p1 = np.zeros(m, np.float32), np.zeros(m, np.uint32)
    L.bvhsim_coop(scene.h, fp(rays), m, 0, 192, up(rd), up(nt), up(ms), fp(t1), up(p1))
    assert (p0 != p1).sum() <= 1e-4 * m, (p0 != p1).sum()  # same hits (exact-t ties on shared edges aside)
    chain = ns + lv
    long_ = ns > 96
    print("\n%s: %d rays, %.1f %% hit" % (name, m, 100 * (p0 != 0xFFFFFFFF).mean()))
    print("  sequential walk: node visits mean %.1f, p99 %d, p99.9 %d, max %d; dependent steps (nodes + leaves) max %d" % (
        ns.mean(), np.quantile(ns, 0.99), np.quantile(ns, 0.999), ns.max(), chain.max()))
    print("  rays over the budget: %d (%.2f %%), holding %.1f %% of all node visits" % (long_.sum(), 100 * long_.mean(), 100 * ns[long_].sum() / ns.sum()))
    if long_.any():
        print("  those rays, sequential: dependent steps mean %.0f, max %d" % (chain[long_].mean(), chain[long_].max()))
        print("  those rays, one warp each from the root: rounds mean %.1f, max %d (%.1fx shorter chain); node tests %.2fx the sequential "
              "count; largest stack %d entries" % (rd[long_].mean(), rd[long_].max(), chain[long_].mean() / rd[long_].mean(),
                                                   nt[long_].sum() / ns[long_].sum(), ms[long_].max()))
    print("  all rays through the cooperative schedule: rounds mean %.1f, max %d; largest stack %d" % (rd.mean(), rd.max(), ms.max()))
