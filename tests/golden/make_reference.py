"""Converged reference renders for the image-level parity bar (relMSE against a converged render, >= 16 k spp
at C1). Run once in the build container (CPU only, ~30 min on 8 cores); the result is committed as a fixture under
tests/golden/ because the GPU box has neither the time nor the reason to re-render it.

usage: make_reference.py [c1|c2|c3] [ref_spp] [probe_spp] [threads] [oracle|reference] [keep-ref]
  oracle     the images are rendered by the oracle port (oracle/oracle_pt.cpp)
  reference  the images are rendered by the REFERENCE ITSELF: ProgressiveMonteCarloIntegrator::render of the libraries that
             oracle/Makefile.ref compiles from /root/reference, with the replay sampler of oracle/ref_harness handing it the
             counter-based stream (seed, pixel, sample) -- so `probe` holds the reference's own image of sample indices
             [0, probe_spp), the very samples the CUDA path renders in tests/test_gpu_image.py
  keep-ref   keep the `ref` image of the existing fixture, re-render only the probe

Writes tests/golden/ref_<name>.npz with
  ref        (H, W, 3) float32  developed image, samples [REF_FIRST, REF_FIRST + ref_spp)  (unguided path tracer, maxDepth 8)
  probe      (H, W, 3) float16  the oracle's own image at probe_spp, samples [0, probe_spp) -- disjoint from ref
  probe_relmse                  relMSE(probe, ref): the noise level an exact implementation has at probe_spp
  meta                          json: scene, sizes, spp, sample ranges, seconds, threads
The GPU test renders samples [0, probe_spp) as well and must reach relMSE <= 1.5 x probe_relmse (tests/test_gpu_image.py).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from b200pg import api  # noqa: E402
from oracle_lib import Oracle, develop  # noqa: E402

REF_FIRST = 1_000_000


def relmse(img, ref):
    """mean over pixels of (I-R)^2 / (R^2 + 1e-3) on developed linear RGB, 0.1 % highest-error pixels discarded"""
    e = ((img.astype(np.float64) - ref) ** 2 / (ref.astype(np.float64) ** 2 + 1e-3)).mean(2).ravel()
    e.sort()
    return float(e[: int(len(e) * 0.999)].mean())


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c1"
    ref_spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    probe_spp = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    threads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    impl = sys.argv[5] if len(sys.argv) > 5 else "oracle"
    keep_ref = len(sys.argv) > 6 and sys.argv[6] == "keep-ref"
    out = os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name)
    if name == "c1":
        sb = pkg.scenes.cornell_box(512, 512, spp=64)
    elif name == "c2":
        sb = pkg.scenes.cornell_caustic(512, 512, spp=64)
    elif name == "c3":  # the medium scene at a quarter of the bench size, 64^3 grid (ProgressiveVolumetricPathTracer)
        sb = pkg.scenes.cornell_medium(256, 256, spp=64, res=64)
    else:
        raise SystemExit("unknown scene")
    p = api.default_params()
    p.max_depth = 8
    p.volumetric = 1 if name == "c3" else 0
    if impl == "reference":
        import ref_lib

        rsc = ref_lib.RefScene(sb)

        def render(first, n):
            return rsc.render(p, first, n, nthreads=threads)[0]
    else:
        sc = Oracle().scene(sb)

        def render(first, n):
            return sc.render(p, first, n, nthreads=threads)[0]
    t0 = time.time()
    acc = np.zeros((sb.height, sb.width, 5), np.float64)  # float32 accumulation over 16 k spp would lose digits
    chunk = 64
    old = dict(np.load(out)) if keep_ref else None
    old_meta = json.loads(str(old["meta"])) if keep_ref else {}
    if keep_ref:
        ref, ref_spp = old["ref"], old_meta["ref_spp"]
    else:
        for s in range(0, ref_spp, chunk):
            acc += render(REF_FIRST + s, min(chunk, ref_spp - s))
            if (s // chunk) % 16 == 0:
                print("ref %d / %d spp, %.0f s" % (s + chunk, ref_spp, time.time() - t0), file=sys.stderr, flush=True)
        ref = develop(acc.astype(np.float64)).astype(np.float32)
    acc[:] = 0
    for s in range(0, probe_spp, chunk):
        acc += render(s, min(chunk, probe_spp - s))
    probe = develop(acc.astype(np.float64)).astype(np.float32)
    meta = dict(scene=name, width=sb.width, height=sb.height, ref_spp=ref_spp, ref_first_sample=REF_FIRST, probe_spp=probe_spp,
                probe_first_sample=0, max_depth=8, seconds=time.time() - t0, threads=threads or os.cpu_count(),
                relmse="mean((I-R)^2/(R^2+1e-3)), 0.1% highest pixels trimmed", probe_rendered_by=impl,
                ref_rendered_by=old_meta.get("ref_rendered_by", "oracle") if keep_ref else impl)
    np.savez_compressed(out, ref=ref, probe=probe.astype(np.float16), probe_relmse=np.float64(relmse(probe, ref)),
                        meta=np.array(json.dumps(meta)))
    print("wrote", out, "probe relMSE", relmse(probe, ref), meta, file=sys.stderr)


if __name__ == "__main__":
    main()
