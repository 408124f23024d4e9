#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native path tracer (driver contract in the task prompt).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    the reference algorithm on the host cores: the guided training
                                                           iteration with the oracle port (the snapshot has no guided
                                                           integrator to run), plus `reference_unguided`: the reference
                                                           ITSELF (oracle/_ref, compiled from its sources) rendering the
                                                           scene with its stock progressivepath next to the port

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


DEFAULT_LANES = 2  # concurrent wavefront sub-batches per progression (the library's default)


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
        try:  # in-process NVML: a query takes well under a millisecond, so even a 0.1-s timed region holds several samples
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))))
                r = int(get_reasons(h))
                for b, n in bits.items():
                    if r & b:
                        self.reasons.add(n)
                time.sleep(0.02)
            return
        except Exception:
            pass  # fall back to the nvidia-smi command line (B200_PROFILING.md's clocks line)
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
    algorithm, pinned sample by sample to the compiled reference; the reference itself has no guided integrator): one progression over a band of rows that records
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


def _reference_unguided_here(pkg, sb, cores, budget_s=8.0):
    """The REFERENCE ITSELF (its libraries and plugins compiled from /root/reference into oracle/_ref by oracle/Makefile.ref,
    driven through oracle/ref_harness: test infrastructure) rendering the workload with its stock configuration --
    `progressivepath` + `independent` sampler, Scene::preprocess + Scene::render on `cores` LocalWorker threads -- next to the
    oracle port rendering the same thing unguided. The reference snapshot has no guided integrator (the guiding
    library is an external dependency that is not in the tree), so the guided CPU baseline stays the port; this leg calibrates
    the port against the real thing on the part both can run. Returns None when oracle/_ref is not there."""
    try:
        import ref_lib

        if not ref_lib.available():
            return {"value": None, "why": "oracle/_ref not built (needs /root/reference at build time)"}
        if sum(s.get("T").shape[0] for s in sb.shapes if s.get("T") is not None) > 2_000_000:
            return {"value": None, "why": "skipped: the reference's SAH kd-tree build over this mesh takes minutes"}
        from oracle_lib import Oracle

        p = pkg._abi.default_params()
        p.max_depth = 8
        p.volumetric = 1 if sb.media else 0
        spp = 1
        rs = ref_lib.RefScene(sb)
        # one untimed warm-up progression (the kd-tree build and the per-pixel sampler allocation are not timed either), then
        # progressions until the budget is spent
        _, t = rs.render(p, 0, spp, nthreads=cores, independent=True, want_film=False, repeat=-int(1e3 * budget_s))
        reps = rs.spp_done // spp
        n = float(sb.width * sb.height * spp * reps)
        osc = Oracle().scene(sb)
        osc.render(p, 0, spp, nthreads=cores)
        to = no = 0.0
        oreps = 0
        while oreps < 2 or (to < budget_s / 2 and oreps < 32):
            _, st = osc.render(p, oreps * spp, spp, nthreads=cores)
            to += st["seconds"]
            no += st["paths"]
            oreps += 1
        return {"value": n / t / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": "reference",
                "sample": "%d unguided progressions of the whole %dx%d image at %d spp through Scene::render (%.1f s); g++ -O2 build "
                          "without the reference's -march=native -funsafe-math-optimizations" % (reps, sb.width, sb.height, spp, t),
                "port_unguided": {"value": no / to / 1e6, "unit": "Mpaths/s", "kind": "port",
                                  "sample": "%d progressions, same scene and parameters (%.1f s)" % (oreps, to)}}
    except Exception as ex:
        return {"value": None, "why": "failed: %s" % ex}


def reference_unguided(workload_name, cores, budget_s=8.0):
    """Runs _reference_unguided_here in a child process (`bench.py --impl reference-unguided`): the reference libraries bring
    their own thread / scheduler / logger singletons, and nothing they do may take the measuring process down."""
    try:
        env = dict(os.environ, OMP_NUM_THREADS=str(cores))
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference-unguided", "--workload", workload_name,
                            "--ref-seconds", str(budget_s)], capture_output=True, text=True, timeout=60 + 20 * budget_s, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"value": None, "why": "child exited %d: %s" % (r.returncode, (r.stderr or "").strip()[-200:])}
    except Exception as ex:
        return {"value": None, "why": "failed: %s" % ex}


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
        # same keys as the GPU arm's config (the driver compares them); every step renders rows 0..`rows` of the image (see
        # cpu_baseline.sample), the throughput figure does not depend on the band height
        "config": {"workload": desc, "spp_per_step": spp, "paths_per_step_per_gpu": sb.width * sb.height * spp,
                   "pretrain_iterations": args.pretrain if guided else 0,
                   "l2": "n/a (host CPU); each step is a bounded band of %d rows of the image" % rows,
                   "guiding": ("training iteration per step: K=16 vMF lobes/cell, %d EM iterations" % args.em_iters) if guided else "off",
                   "parallelism": "%d host threads, 32x32 tiles (imageproc.cpp:27-78)" % cores},
        "mrays_per_sec": rays / t / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample,
                         "kd_traversal_per_ray": cpu.kd_bytes_per_ray(),
                         "note": "kind = port because the step is a GUIDED training iteration and the reference snapshot contains no guided "
                                 "integrator (SURVEY.md F1); the reference itself, compiled from its sources into oracle/_ref, renders the "
                                 "same scene unguided in `reference_unguided`, next to the port doing the same"},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_unguided": reference_unguided(args.workload, cores, args.ref_seconds),
    }
    emit(line)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cornell_caustic_1024")
    ap.add_argument("--spp-per-step", type=int, default=0, help="0 = the workload's default (4; 1 for the 2048^2 mesh)")
    ap.add_argument("--em-iters", type=int, default=4)
    ap.add_argument("--pretrain", type=int, default=12, help="untimed training iterations before warm-up (steady-state field)")
    ap.add_argument("--no-guiding", action="store_true")
    ap.add_argument("--guided-distance", action="store_true", help="medium workloads: guided free-flight sampling")
    ap.add_argument("--nccl-allreduce", action="store_true", help="sum EM statistics with torch.distributed/NCCL instead of the fused peer-memory kernel")
    ap.add_argument("--max-cell-samples", type=int, default=32768, help="spatial split threshold of the guiding field (experiments: a smaller value grows the larger field of a multi-GPU job on one GPU)")
    ap.add_argument("--sort-bounces", type=int, default=-1, help="coherence sort of the shade queue by guiding cell on bounces 1..n (0 = off, -1 = library default)")
    ap.add_argument("--lanes", type=int, default=-1, help="concurrent wavefront sub-batches per progression (-1 = library default)")
    ap.add_argument("--ref-seconds", type=float, default=3.0)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-CPU leg (profiling runs)")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short C3 / C4 runs of the `workloads` block")
    ap.add_argument("--workloads", default="medium_1024,mesh_10m", help="extra BASELINE configs measured (briefly) next to the headline, N = 1 only")
    args = ap.parse_args(argv)
    args.warmup = max(args.warmup, 0)
    if args.spp_per_step <= 0:
        args.spp_per_step = 1 if args.workload == "mesh_10m" else 4
    return args


class Ctx:
    """torch.distributed plumbing shared by all measurements of one process."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            # NCCL_DEBUG=VERSION (set in some images) makes NCCL print its version banner to STDOUT, next to the one JSON line
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_bytes(self, b):
        """all-gather of a fixed-size byte string -> list over ranks"""
        torch, dist = self.torch, self.dist
        mine = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        out = [torch.empty(len(b), dtype=torch.uint8, device="cuda") for _ in range(self.world)]
        dist.all_gather(out, mine)
        return [bytes(g.cpu().numpy().tobytes()) for g in out]


def traffic_table():
    try:  # per-launch DRAM bytes from the committed ncu --set full captures (profiles/)
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def measure(ctx, args, headline):
    """One workload: pre-train, warm up, time `steps` guided training iterations (value), the same through the C-ABI with host
    buffers (e2e), a one-lane profiling pass for the per-kernel roofline figures, and (rank 0, N = 1) the CPU baseline."""
    import hashlib

    torch, dist, rank, world, local = ctx.torch, ctx.dist, ctx.rank, ctx.world, ctx.local
    pkg, sb, desc = workload(args.workload)
    from b200pg import api

    scene = api.Scene.from_builder(sb)
    p = guided_params(pkg, args)
    guided = bool(p.guiding)
    integ = api.Integrator(scene, p, device=local)
    if args.sort_bounces >= 0:
        integ.set_option("sort_bounces", args.sort_bounces)
    lanes = args.lanes if args.lanes > 0 else int(os.environ.get("B200PG_LANES", DEFAULT_LANES))
    overlap = 0 if os.environ.get("B200PG_OVERLAP_SHADOW") == "0" else 1
    integ.set_option("lanes", lanes)
    integ.set_option("overlap_shadow", overlap)
    spp = args.spp_per_step
    npix = sb.width * sb.height
    barrier = ctx.barrier

    def wrap(ptr, n):
        class _W:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_W(), device="cuda")

    def allreduce_stats(ptr, n):  # EM sufficient statistics: sum over ranks (NCCL over NVLink)
        dist.all_reduce(wrap(ptr, n), op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()

    # ---- multi-GPU plumbing: torch.distributed only carries the 64-byte CUDA IPC handles (exchange blocks, films); the
    # per-iteration sum of the EM statistics then happens inside the library's M-step kernel over NVLink peer memory, and
    # previews of the job's film are merged on rank 0's device
    if world > 1:
        if guided and not args.nccl_allreduce:
            integ.comm_connect(rank, world, b"".join(ctx.gather_bytes(integ.comm_local_handle())))
        film_handles = b"".join(ctx.gather_bytes(integ.film_ipc_handle()))
        if rank == 0:
            integ.film_peers_connect(rank, world, film_handles)

    # ---- sample batches are split per GPU: rank r renders sample indices r*spp.. of every step (weak scaling)
    def step(k, rows=None):
        if guided:
            integ.guiding_mode(True, k > 0)
        integ.progression((k * world + rank) * spp, spp, rows=rows)
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
    # ---- timed region: wall clock between two barrier + synchronize brackets (everything the step does on the host and
    # the device); the CUDA-event time of the progressions + training updates is reported next to it as device_total
    t0s = integ.stage_times()
    barrier()
    t_start = time.perf_counter()
    for k in range(args.steps):
        step(base + k)
    barrier()
    wall = time.perf_counter() - t_start
    s1 = integ.stats()
    t1s = integ.stage_times()
    device_total = ctx.max_over_ranks((s1["seconds_total"] - s0["seconds_total"]) + (t1s["train"]["seconds"] - t0s["train"]["seconds"]))
    paths = (s1["paths"] - s0["paths"]) * world
    rays = (s1["normal_rays"] - s0["normal_rays"] + s1["shadow_rays"] - s0["shadow_rays"]) * world
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    wall = ctx.max_over_ranks(wall)
    value = paths / wall / 1e6
    base += args.steps

    # ---- multi-GPU correctness, visible to whoever reads the line: the replicated fields must be bit-identical after the
    # timed updates, and the job's merged film must hold exactly the samples all ranks rendered
    multi = None
    if world > 1:
        h = hashlib.sha1(np.ascontiguousarray(integ.field_snapshot()).tobytes()).digest()[:16] if guided else b"\0" * 16
        hashes = ctx.gather_bytes(h)
        barrier()
        own = integ.film()  # (H, W, 5) of this rank
        w_sum = ctx.sum_over_ranks(float(own[..., 4].astype(np.float64).sum()))
        rgb_sum = ctx.sum_over_ranks(float(own[..., :3].astype(np.float64).sum()))
        multi = {"field_identical_on_all_ranks": bool(all(x == hashes[0] for x in hashes)) if guided else None,
                 "field_sha1_16": hashes[0].hex() if guided else None}
        if rank == 0:
            merged = np.empty_like(own)
            pinned = torch.empty(own.shape, dtype=torch.float32, pin_memory=True).numpy()
            integ.film_async(pinned)
            integ.film_wait()
            merged[:] = pinned
            mw, mrgb = float(merged[..., 4].astype(np.float64).sum()), float(merged[..., :3].astype(np.float64).sum())
            multi.update({"merged_film_weight_sum": mw, "sum_of_rank_weight_sums": w_sum,
                          "merged_film_matches_rank_sum": bool(abs(mw - w_sum) <= 1e-5 * w_sum and abs(mrgb - rgb_sum) <= 1e-4 * abs(rgb_sum)),
                          "samples_in_merged_film": mw, "camera_samples_rendered": float(integ.stats()["paths"]) * world,
                          "note": "a sample's filter weights sum to 1 inside the image (less at the border): weight sum / samples "
                                  "must sit just below 1"})
            multi["weight_sum_over_samples"] = mw / max(multi["camera_samples_rendered"], 1.0)
            assert multi["merged_film_matches_rank_sum"], "the merged film is not the sum of the ranks' films"
            assert 0.97 <= multi["weight_sum_over_samples"] <= 1.0001, "merged film does not hold the samples the ranks rendered"
        barrier()
        assert multi["field_identical_on_all_ranks"] in (True, None), "replicated guiding fields diverged across ranks"

    # ---- roofline of the dominant kernels: ONE-LANE profiling pass (no overlapping streams), same steps, so that every
    # kernel's CUDA-event span is its own; algorithmic bytes from the step's counters
    roof = None
    e2e = None
    cpu = None
    prof_steps = max(2, min(args.steps, 4))
    integ.set_option("lanes", 1)
    integ.set_option("overlap_shadow", 0)
    step(base)  # untimed: the first one-lane step re-sizes the lane's queues
    base += 1
    barrier()
    p0s, ps0 = integ.stage_times(), integ.stats()
    for k in range(prof_steps):
        step(base + k)
    barrier()
    p1s, ps1 = integ.stage_times(), integ.stats()
    base += prof_steps
    tr_sec = p1s["trace"]["seconds"] - p0s["trace"]["seconds"]
    tr_n = p1s["trace"]["launches"] - p0s["trace"]["launches"]
    sh_sec = p1s["shade"]["seconds"] - p0s["shade"]["seconds"]
    sd_sec = p1s["shadow"]["seconds"] - p0s["shadow"]["seconds"]
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
        step_paths = (ps1["paths"] - ps0["paths"]) / prof_steps
        step_nrays = (ps1["normal_rays"] - ps0["normal_rays"]) / prof_steps
        step_srays = (ps1["shadow_rays"] - ps0["shadow_rays"]) / prof_steps
        step_train = (ps1["train_samples"] - ps0["train_samples"]) / prof_steps if guided else 0.0
        shade_bytes = 108.0 * step_nrays + 92.0 * max(step_nrays - step_paths, 0.0) + 52.0 * step_srays + 20.0 * step_paths \
            + 180.0 * step_train
        tn_sec = p1s["train"]["seconds"] - p0s["train"]["seconds"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        sh_n = p1s["shade"]["launches"] - p0s["shade"]["launches"]
        traffic = traffic_table().get(args.workload, {})
        one_lane_total = (ps1["seconds_total"] - ps0["seconds_total"]) + tn_sec
        stage = {"trace": tr_sec / prof_steps, "shade": sh_sec / prof_steps, "shadow": sd_sec / prof_steps,
                 "film": (p1s["film"]["seconds"] - p0s["film"]["seconds"]) / prof_steps, "train": tn_sec / prof_steps,
                 "one_lane_step": one_lane_total / prof_steps, "timed_step_device": device_total / args.steps,
                 "timed_step_host_wall": wall / args.steps,
                 "what": "seconds per step; the stage figures come from a %d-step pass with ONE lane and no stream overlap (each "
                         "kernel's CUDA-event span is exclusive), the timed steps run the library default" % prof_steps}
        ach_shade = shade_bytes * prof_steps / max(sh_sec, 1e-9) / 1e9
        ach_trav = trav_bytes * prof_steps / max(trace_all, 1e-9) / 1e9
        n_prims = sum(int(np.asarray(sh["T"]).shape[0]) if isinstance(sh, dict) and sh.get("T") is not None else 1 for sh in sb.shapes)
        scene_bytes = n_prims * (48 + 96 + 32)  # primitive + shading records + ~half a 64-byte BVH node per primitive
        scene_in_cache = scene_bytes < (32 << 20)
        trav_traffic = traffic.get("trace_dram_bytes_per_step")
        roof = {"bound": "hbm", "achieved": ach_shade, "peak": peak, "unit": "GB/s", "frac": ach_shade / peak,
                "traffic": traffic.get("k_shade_bytes_per_launch"),
                "kernel": "k_shade (intersection fill, NEE, BSDF / guided sampling, queue compaction, training records)",
                "share_of_step": sh_sec / max(one_lane_total, 1e-9),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": shade_bytes * prof_steps / max(sh_n, 1),
                "avg_launch_ms": 1e3 * sh_sec / max(sh_n, 1),
                "stage_seconds": stage,
                "traversal": {"kernel": "k_trace / k_trace_spec + k_shadow_spec (BVH traversal)", "achieved": ach_trav, "frac": ach_trav / peak,
                              "algorithmic_bytes_per_step": trav_bytes,
                              "traffic": trav_traffic,
                              "traffic_frac_of_peak": (trav_traffic / max(trace_all / prof_steps, 1e-9) / 1e9 / peak) if trav_traffic else None,
                              "queue_only_gbs": (48.0 * nrays + 52.0 * srays) * prof_steps / max(trace_all, 1e-9) / 1e9,
                              "per_ray": {"nodes": nodes / max(nrays + srays, 1), "prims": prims / max(nrays + srays, 1)},
                              "avg_launch_ms": 1e3 * tr_sec / max(tr_n, 1),
                              "note": ("the BVH of this small scene is L1-resident, so node / primitive bytes never reach HBM and the "
                                       "algorithmic figure may exceed the HBM peak; queue_only_gbs counts the ray / hit records that do "
                                       "stream through HBM") if scene_in_cache else
                                      ("BVH nodes + primitive records exceed the 126 MB L2: `traffic` = ncu dram__bytes (read + write) of "
                                       "the traversal kernels of one step (profiles/), `achieved` = algorithmic bytes / time")}}
    integ.set_option("lanes", lanes)
    integ.set_option("overlap_shadow", overlap)

    # ---- extra figure: rendering with the trained field, no recording and no training update (what a long guided render
    # does once the training progressions are over); same brackets as the main figure
    render_only = None
    if guided and headline:
        integ.guiding_mode(False, True)
        barrier()
        r0 = integ.stats()
        tr0 = time.perf_counter()
        for k in range(args.steps):
            integ.progression((20_000_000 + k * world + rank) * spp, spp)
        barrier()
        rdt = ctx.max_over_ranks(time.perf_counter() - tr0)
        r1 = integ.stats()
        render_only = {"value": (r1["paths"] - r0["paths"]) * world / rdt / 1e6, "unit": "Mpaths/s", "ms_per_step": 1e3 * rdt / args.steps,
                       "what": "guided rendering with the trained field, no training"}

    # ---- strong scaling extra (N > 1): the SAME step as one GPU renders (spp samples per pixel over the whole image), its rows
    # split into N bands, one per rank; training statistics still summed over the ranks
    strong = None
    if world > 1 and headline:
        rows_per = (sb.height // world + 3) // 4 * 4
        r_lo, r_hi = min(sb.height, rank * rows_per), min(sb.height, (rank + 1) * rows_per)
        if rank == world - 1:
            r_hi = sb.height

        def strong_step(k):
            if guided:
                integ.guiding_mode(True, True)
            if r_hi > r_lo:
                integ.progression((30_000_000 + k) * spp, spp, rows=(r_lo, r_hi))
            if guided:
                integ.train_fused(args.em_iters)
        for k in range(3):
            strong_step(k)
        barrier()
        q0 = integ.stats()
        ts0 = time.perf_counter()
        for k in range(args.steps):
            strong_step(3 + k)
        barrier()
        sdt = ctx.max_over_ranks(time.perf_counter() - ts0)
        q1 = integ.stats()
        sp = ctx.sum_over_ranks(float(q1["paths"] - q0["paths"]))
        strong = {"value": sp / sdt / 1e6, "unit": "Mpaths/s", "ms_per_step": 1e3 * sdt / args.steps, "scaling": "strong",
                  "what": "the single-GPU step (%d spp over the whole %dx%d image) with its rows split into %d bands" % (spp, sb.width, sb.height, world)}

    # ---- end-to-end through the C-ABI with host buffers: scene H2D + render + film D2H every step. With N > 1 the job's
    # film is merged on rank 0's device over NVLink and copied to the host ONCE per step (rank 0), not once per rank.
    barrier()
    host_films = [torch.empty((sb.height, sb.width, 5), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
    if rank == 0:
        integ.film_async(host_films[0])
        integ.film_wait()
    barrier()
    e0 = integ.stats()
    h2d = d2h = 0
    te = time.perf_counter()
    for k in range(args.steps):
        h2d = integ.scene_upload()
        step(base + k)
        # the step's result: a snapshot of the (merged) film, copied device->host on a second stream while the next step
        # renders (double-buffered pinned host arrays; the last snapshot is awaited inside the timed region)
        if rank == 0:
            integ.film_async(host_films[k & 1])
            d2h = host_films[0].nbytes
    if rank == 0:
        integ.film_wait()
    barrier()
    e_elapsed = ctx.max_over_ranks(time.perf_counter() - te)
    t_e2e_end = time.perf_counter()
    clocks.stop_flag = True
    e1 = integ.stats()
    e2e = {"value": (e1["paths"] - e0["paths"]) * world / e_elapsed / 1e6, "unit": "Mpaths/s",
           "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h),
           "what": "scene re-sent host->device on every rank and the job's film read back device->host every step"
                   + (" (merged over NVLink on rank 0's device, one copy for the whole job)" if world > 1 else "")}
    base += args.steps
    if h2d > (64 << 20):  # large scenes: the real API keeps the scene resident across progressions -- report that, too
        barrier()
        f0 = integ.stats()
        tf = time.perf_counter()
        for k in range(args.steps):
            step(base + k)
            if rank == 0:
                integ.film_async(host_films[k & 1])
        if rank == 0:
            integ.film_wait()
        barrier()
        f_el = ctx.max_over_ranks(time.perf_counter() - tf)
        f1 = integ.stats()
        e2e["scene_resident"] = {"value": (f1["paths"] - f0["paths"]) * world / f_el / 1e6, "unit": "Mpaths/s",
                                 "what": "same, scene uploaded once (as b200pg_render does): per step only the film travels"}
        base += args.steps

    line = None
    if rank == 0:
        # ---- CPU baseline: the same guided training iteration with the oracle (port of the reference algorithm) on all
        # host threads, starting from the field the GPU has trained so far, on a bounded band of rows
        try:
            if world > 1:
                raise RuntimeError("reported at N=1 only")
            if args.no_cpu_baseline:
                raise RuntimeError("skipped (--no-cpu-baseline)")
            ncores = len(os.sched_getaffinity(0))  # torchrun sets OMP_NUM_THREADS=1: ask for all host cores explicitly
            tb = time.perf_counter()
            cb = CpuGuidedStep(pkg, sb, p, guided, args.em_iters, ncores)
            build_s = time.perf_counter() - tb
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
                   "sample": "%d guided training iterations over rows 0..%d of the %dx%d image at %d spp (%.1f s of CPU work; kd-tree "
                             "build %.1f s not counted)" % (nrep, rows, sb.width, sb.height, spp, t, build_s) if guided else
                             "%d unguided progressions over rows 0..%d at %d spp (%.1f s)" % (nrep, rows, spp, t),
                   "mrays_per_sec": nr2 / t / 1e6, "kd_traversal_per_ray": cb.kd_bytes_per_ray()}
            if headline:
                cpu["reference_unguided"] = reference_unguided(args.workload, ncores, args.cpu_baseline_seconds / 2)
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
            "lanes": lanes,
        }
        if multi is not None:
            line["multi_gpu_checks"] = multi
        if strong is not None:
            line["strong_scaling"] = strong
        if cpu and cpu.get("value"):
            line["vs_cpu_baseline"] = {"value_ratio": value / cpu["value"], "e2e_ratio": e2e["value"] / cpu["value"]}
            if "scene_resident" in e2e:  # the CPU leg does not rebuild its kd-tree per step either (build time reported in `sample`)
                line["vs_cpu_baseline"]["e2e_scene_resident_ratio"] = e2e["scene_resident"]["value"] / cpu["value"]
    integ.close()
    scene.close()
    return line


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # Library chatter (NCCL prints its version banner to STDOUT at NCCL_DEBUG >= VERSION, which some images set) must not sit
    # next to the JSON line: file descriptor 1 is pointed at stderr for the whole run, the line goes to a saved duplicate.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-unguided":  # child of reference_unguided()
        pkg, sb, _ = workload(args.workload)
        return emit(_reference_unguided_here(pkg, sb, len(os.sched_getaffinity(0)), args.ref_seconds))
    ctx = Ctx()
    line = measure(ctx, args, headline=True)
    # ---- the other BASELINE configs that fit one GPU, measured briefly next to the headline (N = 1 only; the N > 1 runs of the
    # driver's scaling sweep stay short): C3 heterogeneous medium with guided distance sampling, C4 10 M-triangle mesh
    if ctx.world == 1 and not args.no_workloads and args.workload == "cornell_caustic_1024" and line is not None:
        extra = {}
        for name in [w for w in args.workloads.split(",") if w]:
            a = parse_args([])
            a.workload = name
            a.steps, a.warmup, a.pretrain = 6, 3, 8
            a.spp_per_step = 1 if name == "mesh_10m" else 4
            a.guided_distance = name == "medium_1024"
            a.cpu_baseline_seconds = 6.0
            a.no_cpu_baseline = args.no_cpu_baseline
            a.lanes = args.lanes
            try:
                w = measure(ctx, a, headline=False)
                extra[name] = {k: w[k] for k in ("value", "unit", "ms_per_step", "steps", "config", "mrays_per_sec", "e2e", "roofline",
                                                 "cpu_baseline", "vs_cpu_baseline", "clocks") if k in w}
            except Exception as ex:
                extra[name] = {"error": str(ex)}
        line["workloads"] = extra
    if ctx.rank == 0:
        emit(line)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
