#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native path tracer (driver contract in the task prompt).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    the reference algorithm on the host cores
                                                           (oracle restatement; the reference binary cannot
                                                           be built here)

A "step" = one guiding TRAINING ITERATION of the guided path tracer: one progression (``--spp-per-step`` samples per
pixel, recording path-vertex samples and sampling from the current field) over the whole image, followed by the
training update (radix-sort binning, ``--em-iters`` weighted-EM iterations, spatial split). With N > 1 GPUs the
per-cell EM sufficient statistics are summed over the ranks once per EM iteration -- the only exchange on the data path --
inside the library's M-step kernel over NVLink peer memory (CUDA IPC; torch.distributed only carries the 64-byte handles
and the final film reduce); ``--nccl-allreduce`` runs the NCCL allreduce variant for comparison.
value = camera paths completed per second (whole job, all ranks), device-timed with CUDA events, scene resident in HBM. e2e = same metric through the C-ABI with host buffers: per step the
compiled scene is re-sent host->device and the film is read back device->host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def workload(name):
    import __graft_entry__ as ge

    pkg = ge.load_package()
    if name == "cornell_caustic_1024":
        return pkg, pkg.scenes.cornell_caustic(1024, 1024, spp=64), "C2: Cornell box, small shielded light + glass cube, 1024x1024, maxDepth 8"
    if name == "medium_1024":
        return pkg, pkg.scenes.cornell_medium(1024, 1024, spp=64, res=256), "C3: Cornell walls + heterogeneous gridvolume medium (256^3, hg g=0.7), 1024x1024, maxDepth 8"
    if name == "mesh_10m":
        return pkg, pkg.scenes.mesh_scene(2048, 2048, spp=16), "C4: 10.0 M-triangle procedural mesh, roughconductor/roughplastic, 2048x2048, maxDepth 8"
    if name == "cornell_512":
        return pkg, pkg.scenes.cornell_box(512, 512, spp=64), "C1: Cornell box 512x512, maxDepth 8"
    raise SystemExit("unknown workload " + name)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.sm_max = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append((time.perf_counter(), float(out[0])))
                self.sm_max = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self, t0, t1):
        inside = [v for (t, v) in self.samples if t0 <= t <= t1]
        window = "timed regions (value + e2e)"
        if not inside:  # nvidia-smi takes ~0.1 s per query: a very short timed region may see no sample
            inside = [v for (_, v) in self.samples]
            window = "warm-up + timed regions (same load)"
        return {"sm_mhz": float(np.median(inside)) if inside else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(inside), "window": window}


class CpuGuidedStep:
    """The same training iteration as the GPU arm, on the host cores, with the oracle (a CPU restatement of the reference
    algorithm; the reference binary cannot be built here): one progression over a band of rows that records
    path-vertex samples and samples from the current field, then the training update (binning, EM iterations, split)."""

    def __init__(self, pkg, sb, p, guided, em_iters, cores):
        from oracle_lib import Oracle

        self.orc = Oracle()
        self.sc = self.orc.scene(sb)
        self.p, self.sb, self.guided, self.em_iters, self.cores = p, sb, guided, em_iters, cores
        self.field = self.orc.field(p.guide_max_components, (0, 0, 0), (1, 1, 1)) if guided else None
        self.sink = self.orc.samples() if guided else None
        self.trained = False
        self.k = 0
        self.kd = dict(kd_nodes=0, kd_indices=0, prim_tests=0, normal_rays=0, shadow_rays=0)

    def load_field(self, words):
        self.field.load(words)
        self.trained = True

    def step(self, rows, spp):
        """returns (seconds, paths, rays)"""
        t0 = time.perf_counter()
        if self.guided:
            self.sink.clear()
        film, st = self.sc.render(self.p, self.k * spp, spp, rows=(0, rows), nthreads=self.cores,
                                  field=self.field if (self.guided and self.trained) else None, sink=self.sink)
        if self.guided:
            self.field.train_sink(self.sink, self.em_iters, float(self.p.guide_max_cell_samples))
            self.trained = True
        self.k += 1
        for key in ("kd_nodes", "kd_indices", "prim_tests"):
            self.kd[key] += st[key]
        self.kd["normal_rays"] += st["normal_rays"]
        self.kd["shadow_rays"] += st["shadow_rays"]
        return time.perf_counter() - t0, st["paths"], st["normal_rays"] + st["shadow_rays"]

    def kd_bytes_per_ray(self):
        """B_ray = 32 (ray) + 16 (hit; 4 for a shadow ray) + 8 n_node + 4 n_idx + 48 n_tri, averaged over the
        exact ray set by the reference's counting traversal (rayIntersectHavranCollectStatistics, sahkdtree3.h:330-429) --
        here: the oracle's restatement of that traversal over everything this object has rendered."""
        k = self.kd
        rays = k["normal_rays"] + k["shadow_rays"]
        if not rays:
            return None
        n_node, n_idx, n_tri = k["kd_nodes"] / rays, k["kd_indices"] / rays, k["prim_tests"] / rays
        hit = (16 * k["normal_rays"] + 4 * k["shadow_rays"]) / rays
        return {"nodes": n_node, "indices": n_idx, "prim_tests": n_tri, "B_ray": 32 + hit + 8 * n_node + 4 * n_idx + 48 * n_tri,
                "structure": "SAH kd-tree of the reference (oracle restatement), counting traversal"}


def guided_params(pkg, args):
    p = pkg._abi.default_params()
    p.max_depth = 8
    p.guiding = 0 if args.no_guiding else 1
    p.guide_max_components = 16
    p.guide_max_cell_samples = args.max_cell_samples
    p.volumetric = 1 if args.workload == "medium_1024" else 0
    p.guided_distance = 1 if (p.volumetric and p.guiding and args.guided_distance) else 0
    return p


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (oracle restatement, all host threads) on the same workload and the
    same step definition (guided training iteration); every step is a bounded band of rows of the image."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg, sb, desc = workload(args.workload)
    p = guided_params(pkg, args)
    guided = bool(p.guiding)
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun sets it to 1
    cpu = CpuGuidedStep(pkg, sb, p, guided, args.em_iters, cores)
    spp = args.spp_per_step
    # bounded sample: a band of rows sized so that one step takes about ref_seconds
    sec, paths, _ = cpu.step(32, spp)
    rate = paths / max(sec, 1e-6)
    rows = int(min(sb.height, max(32, (rate * args.ref_seconds / (sb.width * spp)) // 32 * 32)))
    for _ in range((args.pretrain if guided else 0) + max(args.warmup, 1)):  # same schedule as the GPU arm
        cpu.step(rows, spp)
    t = paths = rays = 0.0
    for k in range(args.steps):
        sec, np_, nr = cpu.step(rows, spp)
        t += sec
        paths += np_
        rays += nr
    value = paths / t / 1e6
    sample = "rows 0..%d of the %dx%d image, %d spp per step, %d steps (%.1f s)" % (rows, sb.width, sb.height, spp, args.steps, t)
    line = {
        "impl": "reference", "metric": "paths_per_sec", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "spp_per_step": spp, "rows": rows, "pretrain_iterations": args.pretrain if guided else 0,
                   "guiding": ("training iteration per step: K=16 vMF lobes/cell, %d EM iterations" % args.em_iters) if guided else "off"},
        "mrays_per_sec": rays / t / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample,
                         "kd_traversal_per_ray": cpu.kd_bytes_per_ray()},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cornell_caustic_1024")
    ap.add_argument("--spp-per-step", type=int, default=4)
    ap.add_argument("--em-iters", type=int, default=4)
    ap.add_argument("--pretrain", type=int, default=12, help="untimed training iterations before warm-up (steady-state field)")
    ap.add_argument("--no-guiding", action="store_true")
    ap.add_argument("--guided-distance", action="store_true", help="medium workloads: guided free-flight sampling")
    ap.add_argument("--nccl-allreduce", action="store_true", help="sum EM statistics with torch.distributed/NCCL instead of the fused peer-memory kernel")
    ap.add_argument("--max-cell-samples", type=int, default=32768, help="spatial split threshold of the guiding field (experiments: a smaller value grows the larger field of a multi-GPU job on one GPU)")
    ap.add_argument("--sort-bounces", type=int, default=-1, help="coherence sort of the shade queue by guiding cell on bounces 1..n (0 = off, -1 = library default)")
    ap.add_argument("--ref-seconds", type=float, default=3.0)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-CPU leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL_DEBUG=VERSION (set in some images) makes NCCL print its version banner to STDOUT, next to the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    pkg, sb, desc = workload(args.workload)
    from b200pg import api

    scene = api.Scene.from_builder(sb)
    p = guided_params(pkg, args)
    guided = bool(p.guiding)
    integ = api.Integrator(scene, p, device=local)
    if args.sort_bounces >= 0:
        integ.set_option("sort_bounces", args.sort_bounces)
    spp = args.spp_per_step
    npix = sb.width * sb.height

    def wrap(ptr, n):
        class _W:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_W(), device="cuda")

    def allreduce_stats(ptr, n):  # EM sufficient statistics: sum over ranks (NCCL over NVLink)
        dist.all_reduce(wrap(ptr, n), op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()

    # ---- multi-GPU plumbing: torch.distributed only carries the 64-byte CUDA IPC handles of the exchange blocks; the
    # per-iteration sum of the EM statistics then happens inside the library's M-step kernel over NVLink peer memory
    if world > 1 and guided and not args.nccl_allreduce:
        mine = torch.frombuffer(bytearray(integ.comm_local_handle()), dtype=torch.uint8).cuda()
        gathered = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(gathered, mine)
        integ.comm_connect(rank, world, b"".join(bytes(g.cpu().numpy().tobytes()) for g in gathered))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- sample batches are split per GPU: rank r renders sample indices r*spp.. of every step (weak scaling)
    def step(k):
        if guided:
            integ.guiding_mode(True, k > 0)
        integ.progression((k * world + rank) * spp, spp)
        if guided:
            if world > 1 and args.nccl_allreduce:  # comparison path: NCCL allreduce between separate E / M kernels
                integ.train(args.em_iters, allreduce_stats)
            else:
                integ.train_fused(args.em_iters)

    clocks = ClockSampler(local)
    if rank == 0:  # one sampler per job: nvidia-smi takes a driver-wide lock, 8 pollers would perturb the run
        clocks.start()
    # untimed setup: bring the guiding field to its steady state (the spatial tree stops growing after ~10 updates), so
    # that the timed steps -- and the e2e steps after them -- all cost the same; the reference arm does the same
    base = 0
    if guided:
        for k in range(args.pretrain):
            step(k)
        base = args.pretrain
    for k in range(max(args.warmup, 3)):
        step(base + k)
    base += max(args.warmup, 3)
    barrier()
    s0 = integ.stats()
    t0s = integ.stage_times()
    # device time of the timed region = seconds_total of the stream (CUDA-event drained) -> use CUDA events via torch on
    # our own stream is not visible to torch; the library brackets every progression with stream syncs and reports
    # per-stage CUDA-event times; the step time below is host wall-clock around fully synchronised progressions.
    barrier()
    t_start = time.perf_counter()
    for k in range(args.steps):
        step(base + k)
    barrier()
    wall = time.perf_counter() - t_start
    s1 = integ.stats()
    t1s = integ.stage_times()
    # device time: CUDA events recorded on the launching stream around every progression (b200pg stats)
    elapsed = (s1["seconds_total"] - s0["seconds_total"]) + (t1s["train"]["seconds"] - t0s["train"]["seconds"])
    if world > 1:
        tt = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt.item())
    paths = (s1["paths"] - s0["paths"]) * world
    rays = (s1["normal_rays"] - s0["normal_rays"] + s1["shadow_rays"] - s0["shadow_rays"]) * world
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    # the step time is the wall clock between two barrier + synchronize brackets (everything the step does on the host
    # and the device); `elapsed` above (sum of the CUDA-event spans) is reported next to it as device_total
    if world > 1:
        tw = torch.tensor([wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw.item())
    value = paths / wall / 1e6

    # ---- roofline of the dominant kernel (closest-hit traversal): algorithmic bytes from a counting pass
    tr_sec = t1s["trace"]["seconds"] - t0s["trace"]["seconds"]
    tr_n = t1s["trace"]["launches"] - t0s["trace"]["launches"]
    sh_sec = t1s["shade"]["seconds"] - t0s["shade"]["seconds"]
    sd_sec = t1s["shadow"]["seconds"] - t0s["shadow"]["seconds"]
    roof = None
    e2e = None
    cpu = None
    if rank == 0:
        integ.set_option("count_traversal", 1)
        if guided:
            integ.guiding_mode(False, True)
        c0 = integ.stats()
        integ.progression(10_000_000, spp)  # same workload, disjoint sample indices, outside the timed region
        c1 = integ.stats()
        integ.set_option("count_traversal", 0)
        nrays = c1["normal_rays"] - c0["normal_rays"]
        srays = c1["shadow_rays"] - c0["shadow_rays"]
        nodes = c1["bvh_nodes_visited"] - c0["bvh_nodes_visited"]
        prims = c1["prims_tested"] - c0["prims_tested"]
        # Traversal: B_ray = 32 (ray in) + 16 (hit out; 4 for shadow rays) + 64 B per BVH node visited + 48 B per
        # primitive test; counters cover closest + shadow rays of one step.
        trav_bytes = 32.0 * (nrays + srays) + 16.0 * nrays + 4.0 * srays + 64.0 * nodes + 48.0 * prims
        trace_all = tr_sec + sd_sec
        # Shade stage: path-state streaming. Per queued path 92 B state + 16 B hit in, 92 B out per surviving path,
        # 52 B per shadow-queue entry, 20 B splat record per finished path; per recorded training vertex 64 B record +
        # 16 B close, and per emitted training sample 64 B read back + 36 B written (DESIGN.md "Data layout").
        step_paths = (s1["paths"] - s0["paths"]) / args.steps
        step_nrays = (s1["normal_rays"] - s0["normal_rays"]) / args.steps
        step_srays = (s1["shadow_rays"] - s0["shadow_rays"]) / args.steps
        step_train = (s1["train_samples"] - s0["train_samples"]) / args.steps if guided else 0.0
        shade_bytes = 108.0 * step_nrays + 92.0 * max(step_nrays - step_paths, 0.0) + 52.0 * step_srays + 20.0 * step_paths \
            + 180.0 * step_train
        tn_sec = t1s["train"]["seconds"] - t0s["train"]["seconds"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        sh_n = t1s["shade"]["launches"] - t0s["shade"]["launches"]
        traffic = None
        try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        stage = {"trace": tr_sec, "shade": sh_sec, "shadow": sd_sec, "film": t1s["film"]["seconds"] - t0s["film"]["seconds"],
                 "train": tn_sec, "device_total": elapsed, "host_wall": wall}
        ach_shade = shade_bytes * args.steps / max(sh_sec, 1e-9) / 1e9
        ach_trav = trav_bytes * args.steps / max(trace_all, 1e-9) / 1e9
        # the kernel with the largest share of the step is the shade stage (profiles/: ncu launch list of this command)
        roof = {"bound": "hbm", "achieved": ach_shade, "peak": peak, "unit": "GB/s", "frac": ach_shade / peak,
                "traffic": (traffic or {}).get("k_shade_bytes_per_launch"),
                "kernel": "k_shade (intersection fill, NEE, BSDF / guided sampling, queue compaction, training records)",
                "share_of_step": sh_sec / max(elapsed, 1e-9),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": shade_bytes * args.steps / max(sh_n, 1),
                "avg_launch_ms": 1e3 * sh_sec / max(sh_n, 1),
                "stage_seconds": stage,
                "traversal": {"kernel": "k_trace + k_shadow (BVH traversal)", "achieved": ach_trav, "frac": ach_trav / peak,
                              "algorithmic_bytes_per_step": trav_bytes,
                              "queue_only_gbs": (48.0 * nrays + 52.0 * srays) * args.steps / max(trace_all, 1e-9) / 1e9,
                              "per_ray": {"nodes": nodes / max(nrays + srays, 1), "prims": prims / max(nrays + srays, 1)},
                              "avg_launch_ms": 1e3 * tr_sec / max(tr_n, 1),
                              "note": "the BVH of this 40-primitive scene is L1-resident, so node/primitive bytes never reach HBM "
                                      "and the algorithmic figure may exceed the HBM peak; queue_only_gbs counts the ray/hit "
                                      "records that do stream through HBM"}}

    # ---- extra figure: rendering with the trained field, no recording and no training update (what a long guided render
    # does once the training progressions are over); same brackets as the main figure
    render_only = None
    if guided:
        integ.guiding_mode(False, True)
        barrier()
        r0 = integ.stats()
        tr0 = time.perf_counter()
        for k in range(args.steps):
            integ.progression((20_000_000 + k * world + rank) * spp, spp)
        barrier()
        rdt = time.perf_counter() - tr0
        if world > 1:
            tt = torch.tensor([rdt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            rdt = float(tt.item())
        r1 = integ.stats()
        render_only = {"value": (r1["paths"] - r0["paths"]) * world / rdt / 1e6, "unit": "Mpaths/s", "ms_per_step": 1e3 * rdt / args.steps,
                       "what": "guided rendering with the trained field, no training"}

    # ---- end-to-end through the C-ABI with host buffers: scene H2D + render + film D2H every step
    barrier()
    host_film = torch.empty((sb.height, sb.width, 5), dtype=torch.float32, pin_memory=True).numpy()  # pinned host buffer
    host_films = [host_film, torch.empty((sb.height, sb.width, 5), dtype=torch.float32, pin_memory=True).numpy()]
    integ.film(out=host_film)
    e0 = integ.stats()
    h2d = d2h = 0
    te = time.perf_counter()
    for k in range(args.steps):
        _t0 = time.perf_counter()
        h2d = integ.scene_upload()
        _t1 = time.perf_counter()
        step(base + args.steps + k)
        _t2 = time.perf_counter()
        # the step's result: a snapshot of the film, copied device->host on a second stream while the next step renders
        # (double-buffered pinned host arrays; the last snapshot is awaited inside the timed region)
        integ.film_async(host_films[k & 1])
        d2h = host_film.nbytes
        if os.environ.get("B200PG_BENCH_DEBUG"):
            print("e2e step", k, "upload %.2f step %.2f film %.2f ms" % (1e3 * (_t1 - _t0), 1e3 * (_t2 - _t1), 1e3 * (time.perf_counter() - _t2)), file=sys.stderr)
    integ.film_wait()
    _tb = time.perf_counter()
    barrier()
    e_elapsed = time.perf_counter() - te
    t_e2e_end = time.perf_counter()
    clocks.stop_flag = True
    if os.environ.get("B200PG_BENCH_DEBUG"):
        print("e2e rank", rank, "loop %.2f ms, closing barrier %.2f ms" % (1e3 * (_tb - te), 1e3 * (time.perf_counter() - _tb)), file=sys.stderr)
    if world > 1:
        tt = torch.tensor([e_elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e_elapsed = float(tt.item())
    e1 = integ.stats()
    e2e = {"value": (e1["paths"] - e0["paths"]) * world / e_elapsed / 1e6, "unit": "Mpaths/s",
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    # ---- multi-GPU: every rank holds a full-size film; one NCCL reduce at the end
    if world > 1:
        ptr, n = integ.film_device_buffer()
        dist.reduce(wrap(ptr, n), dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()

    if rank == 0:
        # ---- CPU baseline: the same guided training iteration with the oracle (port of the reference algorithm) on all
        # host threads, starting from the field the GPU has trained so far, on a bounded band of rows (~10-30 s in total)
        try:
            if world > 1:
                raise RuntimeError("reported at N=1 only")
            if args.no_cpu_baseline:
                raise RuntimeError("skipped (--no-cpu-baseline)")
            ncores = len(os.sched_getaffinity(0))  # torchrun sets OMP_NUM_THREADS=1: ask for all host cores explicitly
            cb = CpuGuidedStep(pkg, sb, p, guided, args.em_iters, ncores)
            if guided:
                cb.load_field(integ.field_snapshot())
            sec, np_, nr = cb.step(32, spp)
            rate = np_ / max(sec, 1e-6)
            rows = int(min(sb.height, max(32, (rate * args.cpu_baseline_seconds / 3 / (sb.width * spp)) // 32 * 32)))
            t = np2 = nr2 = 0.0
            nrep = 0
            while nrep < 3 or (t < args.cpu_baseline_seconds and nrep < 64):
                sec, np_, nr = cb.step(rows, spp)
                t += sec
                np2 += np_
                nr2 += nr
                nrep += 1
            cpu = {"value": np2 / t / 1e6, "unit": "Mpaths/s", "cores": ncores, "kind": "port",
                   "sample": "%d guided training iterations over rows 0..%d of the %dx%d image at %d spp (%.1f s of CPU work)"
                             % (nrep, rows, sb.width, sb.height, spp, t) if guided else
                             "%d unguided progressions over rows 0..%d at %d spp (%.1f s)" % (nrep, rows, spp, t),
                   "mrays_per_sec": nr2 / t / 1e6, "kd_traversal_per_ray": cb.kd_bytes_per_ray()}
        except Exception as ex:  # the oracle is test infrastructure; its absence must not break the product arm
            cpu = {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "port", "sample": "oracle unavailable: %s" % ex}
        line = {
            "metric": "paths_per_sec", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "spp_per_step": spp, "paths_per_step_per_gpu": npix * spp, "pretrain_iterations": args.pretrain if guided else 0,
                       "l2": "wavefront state per step (%.0f MB) exceeds the 126 MB L2" % (npix * spp * 250 / 1e6),
                       "guiding": ("training iteration per step: K=16 vMF lobes/cell, %d EM iterations, %d cells at the end"
                                   % (args.em_iters, s1["guide_cells"])) if guided else "off",
                       "parallelism": ("sample batches split per GPU; EM statistics summed over NVLink peer memory inside the M-step kernel"
                                       if not args.nccl_allreduce else "sample batches split per GPU; NCCL allreduce of EM statistics") if world > 1 else "1 GPU"},
            "mrays_per_sec": rays / wall / 1e6,
            "gpu_launches": int(launches),
            "clocks": clocks.summary(t_start, t_e2e_end),
            "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "render_only": render_only,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
