"""CPU arm of tools/equal_time.py: the oracle (CPU restatement of the reference, all host cores) rendering progressions for a wall-clock
budget with the same training schedule as the GPU arm. Lives under tests/ because it loads the oracle (test infrastructure)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def cpu_arm(sb, p, T, guided, train_passes, W, H, ref, floor, relmse):
    from oracle_lib import Oracle, develop

    orc = Oracle()
    osc = orc.scene(sb)
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    field = orc.field(16, (0, 0, 0), (1, 1, 1)) if guided else None
    sink = orc.samples() if guided else None
    film = np.zeros((H, W, 5), np.float32)
    acc = np.zeros((H, W, 5), np.float64)
    t0 = time.perf_counter()
    k = paths = 0
    trained = False
    while time.perf_counter() - t0 < T:
        film[:] = 0
        training = guided and k < train_passes
        if training:
            sink.clear()
        _, st = osc.render(p, k, 1, film=film, nthreads=cores, field=field if (guided and trained) else None, sink=sink if training else None)
        if training:
            field.train_sink(sink, 4, float(p.guide_max_cell_samples))
            trained = True
        acc += film
        paths += st["paths"]
        k += 1
    el = time.perf_counter() - t0
    r = relmse(develop(acc).astype(np.float32), ref)
    return {"relMSE": r, "relMSE_debiased": max(r - floor, 0.0), "seconds": el, "spp": paths / (W * H), "mpaths_per_s": paths / el / 1e6,
            "cores": cores}
