"""EM statistics exchange microbenchmark (BASELINE.json configs[4]: "guiding-field EM allreduce scaling at
1/2/4/8 GPUs"): cells in {1 k, 8 k, 64 k} x K = 32 lobes (136 floats = 544 B of sufficient statistics per cell).

For every size, on N ranks of one box (python -m torch.distributed.run --nproc-per-node N tools/em_exchange_bench.py):
  fused  = k_mstep_allreduce: cross-GPU barrier + per-cell sum over the ranks' buffers read straight from peer HBM over
           NVLink + M-step, ONE kernel (what b200pg_train runs between EM iterations); its two forms are also timed
           separately: allread (every rank reads every peer's buffer; small fields) and twophase (every rank sums its
           slice of the cells and pushes the sums to all ranks: reduce-scatter + all-gather traffic; large fields)
  local  = the same kernel over the rank's own buffer only (the M-step share, no exchange)
  nccl   = torch.distributed.all_reduce(sum) of the same buffer (what the north star names); the unfused alternative
           costs nccl + local
All times are device times (CUDA events), max over ranks. One JSON line per size on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29731")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from conftest import load_package

    pkg = load_package()
    from b200pg import api

    K = 32
    p = api.default_params()
    p.max_depth, p.guiding, p.guide_max_components = 8, 1, K
    it = api.Integrator(api.Scene.from_builder(pkg.scenes.cornell_caustic(64, 64, spp=1)), p, device=local)
    mine = torch.frombuffer(bytearray(it.comm_local_handle()), dtype=torch.uint8).cuda()
    gathered = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, mine)
    it.comm_connect(rank, world, b"".join(g.cpu().numpy().tobytes() for g in gathered))
    stride = 4 * K + 8
    iters = 50

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for cells in (1024, 8192, 65536):
        res = {}
        for name, mode in (("fused_ms", 0), ("local_mstep_ms", 1), ("allread_ms", 2), ("twophase_ms", 3)):
            dist.barrier()
            torch.cuda.synchronize()
            res[name] = max_over_ranks(it.em_exchange_bench(cells, iters, mode))
        fused, local_ms = res["fused_ms"], res["local_mstep_ms"]
        buf = torch.rand(cells * stride, device="cuda")
        for _ in range(5):
            dist.all_reduce(buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            dist.all_reduce(buf)
        e1.record()
        torch.cuda.synchronize()
        nccl = max_over_ranks(e0.elapsed_time(e1) / iters)
        if rank == 0:
            nbytes = cells * stride * 4
            print(json.dumps({
                "bench": "em_exchange", "n_gpus": world, "cells": cells, "K": K, "stats_bytes_per_rank": nbytes,
                "fused_ms": round(fused, 5), "allread_ms": round(res["allread_ms"], 5), "twophase_ms": round(res["twophase_ms"], 5),
                "local_mstep_ms": round(local_ms, 5), "nccl_allreduce_ms": round(nccl, 5),
                "unfused_ms": round(nccl + local_ms, 5),
                "iters": iters}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
