"""Two-GPU check of the fused cross-GPU statistics sum + M-step (k_mstep_allreduce over NVLink peer memory, CUDA IPC):
two ranks render disjoint sample batches, train through b200pg_train with connected peers, and must end up with
(i) bit-identical fields on both ranks and (ii) the field a single GPU trains from the union of the samples (within the
parity bar of the EM statistics: the sums are the same numbers added in a different order). Needs 2 visible GPUs;
skipped otherwise. torch.distributed (NCCL) only carries the 64-byte IPC handles."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _params(api):
    p = api.default_params()
    p.max_depth, p.guiding, p.guide_max_components, p.guide_max_cell_samples = 8, 1, 16, 6000
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_package

    pkg = load_package()
    from b200pg import api

    sb = pkg.scenes.cornell_caustic(128, 128, spp=8)
    it = api.Integrator(api.Scene.from_builder(sb), _params(api), device=rank)
    mine = torch.frombuffer(bytearray(it.comm_local_handle()), dtype=torch.uint8).cuda()
    gathered = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, mine)
    it.comm_connect(rank, world, b"".join(g.cpu().numpy().tobytes() for g in gathered))
    for k in range(4):
        it.guiding_mode(True, k > 0)
        it.progression(100 * k + 2 * rank, 2)  # rank r renders samples 100k + 2r, 100k + 2r + 1
        it.train_fused(4)
    np.save(os.path.join(out_dir, "field_%d.npy" % rank), it.field_snapshot())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_fused_statistics_sum(tmp_path, pkg):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from b200pg import api

    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    f0, f1 = np.load(tmp_path / "field_0.npy"), np.load(tmp_path / "field_1.npy")
    assert np.array_equal(f0, f1)  # replicated fields stay bit-identical without a broadcast
    # the reduce-scatter + all-gather form of the exchange (large fields) adds the same numbers in the same rank order:
    # forced on through the environment it must reproduce the all-read form's field bit for bit
    os.environ["B200PG_EXCHANGE_FORM"] = "1"
    try:
        mp.spawn(_worker, args=(2, port + 1, str(tmp_path)), nprocs=2, join=True)
    finally:
        del os.environ["B200PG_EXCHANGE_FORM"]
    g0, g1 = np.load(tmp_path / "field_0.npy"), np.load(tmp_path / "field_1.npy")
    assert np.array_equal(g0, g1)
    assert g0.shape == f0.shape and np.array_equal(g0[:8], f0[:8])  # same tree size
    # single GPU, union of the samples: the first update sees exactly the same samples (unguided progression), so the
    # spatial tree after it must be identical and the mixtures agree to the statistics' tolerance
    sb = pkg.scenes.cornell_caustic(128, 128, spp=8)
    it = api.Integrator(api.Scene.from_builder(sb), _params(api))
    it.guiding_mode(True, False)
    it.progression(0, 4)
    it.train_fused(4)
    one = it.field_snapshot()
    nn = int(one[1])
    assert nn >= 3  # the first update split the root cell
    two_first = _first_update_two_rank(tmp_path)
    # same tree: identical topology; the split planes are means of float position sums that the two ranks accumulate in a
    # different order than one GPU does (1e-6 relative)
    assert np.array_equal(one[:8], two_first[:8])
    na, nb = one[8:8 + 4 * nn].reshape(nn, 4), two_first[8:8 + 4 * nn].reshape(nn, 4)
    assert np.array_equal(na[:, [0, 2, 3]], nb[:, [0, 2, 3]])
    np.testing.assert_allclose(nb[:, 1].view(np.float32), na[:, 1].view(np.float32), rtol=2e-6, atol=1e-6)
    o = 8 + 4 * nn + 8 * int(one[2])
    la, lb = one.view(np.float32)[o:].reshape(-1, 12), two_first.view(np.float32)[o:].reshape(-1, 12)
    assert np.abs(la[:, 0] - lb[:, 0]).max() <= 1e-5
    heavy = la[:, 0] > 1e-3
    assert np.abs(la[heavy, 1:4] - lb[heavy, 1:4]).max() <= 5e-5


def _worker_first(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_package

    pkg = load_package()
    from b200pg import api

    sb = pkg.scenes.cornell_caustic(128, 128, spp=8)
    it = api.Integrator(api.Scene.from_builder(sb), _params(api), device=rank)
    mine = torch.frombuffer(bytearray(it.comm_local_handle()), dtype=torch.uint8).cuda()
    gathered = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, mine)
    it.comm_connect(rank, world, b"".join(g.cpu().numpy().tobytes() for g in gathered))
    it.guiding_mode(True, False)
    it.progression(2 * rank, 2)  # samples 0,1 on rank 0 and 2,3 on rank 1 = the single-GPU progression(0, 4)
    it.train_fused(4)
    if rank == 0:
        np.save(os.path.join(out_dir, "first.npy"), it.field_snapshot())
    dist.barrier()
    dist.destroy_process_group()


def _first_update_two_rank(tmp_path):
    import torch.multiprocessing as mp

    port = 29900 + os.getpid() % 90
    mp.spawn(_worker_first, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    return np.load(tmp_path / "first.npy")
