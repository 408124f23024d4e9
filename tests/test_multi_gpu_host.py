"""N > 1 host logic on CPU (gloo, world_size 2): the only data exchanged between ranks are the per-cell EM sufficient
statistics (sum) and the film (sum). Both are checked with the oracle standing in for the per-rank compute:
shard the samples / the sample indices over two ranks, all-reduce, and compare with the unsharded result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_package
    from oracle_lib import Oracle

    pkg = load_package()
    orc = Oracle()
    sb = pkg.scenes.cornell_caustic(32, 32, spp=4)
    sc = orc.scene(sb)
    p = pkg._abi.default_params()
    p.max_depth = 5
    fld = orc.field(8, (-1.1, -0.1, -1.1), (1.1, 2.1, 1.1))
    # --- rank r renders sample indices r, r + world, ... (sample batches split per GPU)
    sink = orc.samples()
    film = np.zeros((32, 32, 5), np.float32)
    for s in range(rank, 4, world):
        sc.render(p, s, 1, film=film, sink=sink, nthreads=1)
    smp = sink.get()
    # --- EM statistics: local E-step, sum over ranks, identical M-step everywhere
    stats = torch.from_numpy(fld.estep(smp).copy())
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    ft = torch.from_numpy(film)
    dist.all_reduce(ft, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, "stats_%d.npy" % rank), stats.numpy())
    np.save(os.path.join(out_dir, "film_%d.npy" % rank), ft.numpy())
    dist.destroy_process_group()


def test_two_rank_statistics_and_film_reduce(tmp_path, pkg, oracle):
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    s0, s1 = np.load(tmp_path / "stats_0.npy"), np.load(tmp_path / "stats_1.npy")
    f0, f1 = np.load(tmp_path / "film_0.npy"), np.load(tmp_path / "film_1.npy")
    assert np.array_equal(s0, s1) and np.array_equal(f0, f1)  # every rank holds the same reduced data
    # single-process reference
    sb = pkg.scenes.cornell_caustic(32, 32, spp=4)
    sc = oracle.scene(sb)
    p = pkg._abi.default_params()
    p.max_depth = 5
    fld = oracle.field(8, (-1.1, -0.1, -1.1), (1.1, 2.1, 1.1))
    sink = oracle.samples()
    film = np.zeros((32, 32, 5), np.float32)
    for s in range(4):
        sc.render(p, s, 1, film=film, sink=sink, nthreads=1)
    whole = fld.estep(sink.get())
    np.testing.assert_allclose(s0, whole, rtol=1e-5, atol=1e-5)
    assert s0[0, -8] == whole[0, -8]  # integer sample counts reduce exactly
    np.testing.assert_allclose(f0, film, rtol=1e-5, atol=1e-5)
