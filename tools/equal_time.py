"""relMSE at equal time (north-star metric) through the product's own time-budget loop:
b200pg_render with maxRenderTime (ProgressiveMonteCarloIntegrator::renderTime, progressiveintegrator.cpp:117-168) on 1..N GPUs
(in-library device list), guided (training progressions inside the budget) against unguided, next to the CPU arm -- the oracle
port of the reference on all host cores -- at the same budgets.

  equal_time.py --scene c2 --size 512 --budgets 1,3,10 [--gpus N] [--cpu]         reference = the converged ORACLE render
                                                                                    (tests/golden/ref_c2.npz, 16 384 spp)
  equal_time.py --scene c2 --size 3840x2160 --budgets 10,30 --ref-spp 8192 ...    reference = unguided GPU render at --ref-spp
        (a 4K oracle reference is 136 G paths = half a day on 16 cores; the GPU estimator is pinned to the oracle's at 512^2 by
         tests/test_gpu_image.py and by the first mode of this script)

relMSE = mean over pixels of (I - R)^2 / (R^2 + 1e-3) on developed linear RGB, 0.1 % highest-error pixels discarded.
One JSON line per budget on stdout.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from b200pg import api  # noqa: E402


def relmse(img, ref):
    e = ((img.astype(np.float64) - ref) ** 2 / (ref.astype(np.float64) ** 2 + 1e-3)).mean(2).ravel()
    e.sort()
    return float(e[: int(len(e) * 0.999)].mean())


def params(guided, seconds=0, spp_per_pass=4, train_passes=16):
    p = api.default_params()
    p.max_depth = 8
    p.samples_per_progression = spp_per_pass
    p.max_render_time = int(seconds)
    p.guiding = 1 if guided else 0
    p.guide_max_components = 16
    p.guide_max_cell_samples = 32768
    p.training_progressions = train_passes if guided else 0
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c2", choices=["c1", "c2"])
    ap.add_argument("--size", default="512")
    ap.add_argument("--budgets", default="1,3,10")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--ref-spp", type=int, default=0, help="0 = the oracle fixture (512^2 only); > 0 = unguided GPU render with that many spp")
    ap.add_argument("--cpu", action="store_true", help="also run the CPU arm (oracle, all host cores) at the same budgets")
    ap.add_argument("--train-passes", type=int, default=16)
    ap.add_argument("--split-levels", default="1", help="comma list: spatial split levels per training update (set_option split_levels); "
                                                        "the first entry is the `guided` arm, further ones are reported as guided_levels<n>")
    args = ap.parse_args()
    W, H = (int(x) for x in args.size.split("x")) if "x" in args.size else (int(args.size), int(args.size))
    budgets = [float(x) for x in args.budgets.split(",")]
    make = pkg.scenes.cornell_box if args.scene == "c1" else pkg.scenes.cornell_caustic
    sb = make(W, H, spp=64)
    scene = api.Scene.from_builder(sb)
    devices = list(range(args.gpus))

    floor = 0.0
    if args.ref_spp > 0:
        # two independent halves (different seeds): their mean is the reference, their difference measures the reference's OWN
        # noise -- E[(A - B)^2] / 4 = var(reference) -- which every relMSE below contains and `relMSE_debiased` subtracts
        half = max(1, args.ref_spp // 2)
        t0 = time.perf_counter()
        halves = []
        for k in range(2):
            sbr = make(W, H, spp=half)
            sbr.seed = sb.seed + 104729 * (k + 1)
            it = api.Integrator(api.Scene.from_builder(sbr), params(False, 0, 64))
            it.render(devices=devices if len(devices) > 1 else None)
            halves.append(it.develop().astype(np.float64))
            it.close()
        ref = (0.5 * (halves[0] + halves[1])).astype(np.float32)
        e = (((halves[0] - halves[1]) ** 2) / (ref.astype(np.float64) ** 2 + 1e-3)).mean(2).ravel()
        e.sort()
        floor = float(e[: int(len(e) * 0.999)].mean()) / 4.0
        ref_desc = "unguided GPU render, 2 x %d spp on %d GPU(s), %.1f s; its own noise (from the two halves) adds %.2e to every relMSE" % (
            half, len(devices), time.perf_counter() - t0, floor)
        seed_shift = 7919
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % args.scene))
        meta = json.loads(str(z["meta"]))
        assert (meta["width"], meta["height"]) == (W, H), "the oracle fixture is %dx%d" % (meta["width"], meta["height"])
        ref = z["ref"].astype(np.float32)
        ref_desc = "converged oracle render, %d spp (tests/golden/ref_%s.npz), samples from index %d" % (meta["ref_spp"], args.scene, meta["ref_first_sample"])
        seed_shift = 0
        # the reference's own noise: what relMSE an EXACT image would still show against it
        floor = float(z["probe_relmse"]) * meta["probe_spp"] / meta["ref_spp"]
        ref_desc += "; its own noise adds %.2e to every relMSE" % floor
    print("reference:", ref_desc, file=sys.stderr)

    sbt = make(W, H, spp=64)
    sbt.seed = sb.seed + seed_shift  # disjoint from a GPU-rendered reference (sample streams are keyed by the seed)

    def gpu_arm(T, guided, levels=1):
        sc = api.Scene.from_builder(sbt)
        it = api.Integrator(sc, params(guided, T, 4, args.train_passes))
        it.progression(0, 1)  # warm the allocations outside the budget, then start from an empty film / field
        it.film_clear()
        it.close()
        it = api.Integrator(sc, params(guided, T, 4, args.train_passes))
        if guided and levels > 1:
            it.set_option("split_levels", levels)
        t0 = time.perf_counter()
        it.render(devices=devices if len(devices) > 1 else None)
        el = time.perf_counter() - t0
        img, st = it.develop(), it.stats()
        r = relmse(img, ref)
        out = {"relMSE": r, "relMSE_debiased": max(r - floor, 0.0), "seconds": el, "spp": st["paths"] / (W * H), "mpaths_per_s": st["paths"] / el / 1e6,
               "cells": st.get("guide_cells", 0), "progressions": st["progressions_done"]}
        it.close()
        return out

    def cpu_arm(T, guided):
        # the CPU arm runs the oracle, which is test infrastructure: it lives under tests/ (tests/manual/equal_time_cpu.py)
        sys.path.insert(0, os.path.join(ROOT, "tests", "manual"))
        from equal_time_cpu import cpu_arm as _cpu

        return _cpu(sbt, params(guided), T, guided, args.train_passes, W, H, ref, floor, relmse)

    for T in budgets:
        levels = [int(x) for x in args.split_levels.split(",")]
        out = {"scene": args.scene, "size": "%dx%d" % (W, H), "budget_s": T, "gpus": args.gpus, "reference": ref_desc,
               "guided": gpu_arm(T, True, levels[0]), "unguided": gpu_arm(T, False)}
        out["guided"]["split_levels"] = levels[0]
        for lv in levels[1:]:
            out["guided_levels%d" % lv] = gpu_arm(T, True, lv)
        out["reference_noise"] = floor
        out["relMSE_ratio_unguided_over_guided"] = out["unguided"]["relMSE"] / out["guided"]["relMSE"]
        if out["guided"]["relMSE_debiased"] > 0:
            out["relMSE_debiased_ratio_unguided_over_guided"] = out["unguided"]["relMSE_debiased"] / out["guided"]["relMSE_debiased"]
        if args.cpu:
            out["cpu_guided"] = cpu_arm(T, True)
            out["cpu_unguided"] = cpu_arm(T, False)
            out["relMSE_ratio_cpu_guided_over_gpu_guided"] = out["cpu_guided"]["relMSE"] / out["guided"]["relMSE"]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
