import os, sys, time
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge
pkg = ge.load_package()
from b200pg import api
import torch
sb = pkg.scenes.cornell_caustic(1024, 1024)
scene = api.Scene.from_builder(sb)
p = api.default_params(); p.max_depth = 8; p.guiding = 1; p.guide_max_components = 16; p.guide_max_cell_samples = 32768
it = api.Integrator(scene, p)
host = torch.empty((1024, 1024, 5), dtype=torch.float32, pin_memory=True).numpy()
for k in range(24):
    t0 = time.perf_counter(); it.scene_upload(); t1 = time.perf_counter()
    it.guiding_mode(True, k > 0); it.progression(4 * k, 4); t2 = time.perf_counter()
    n, c = it.train(4); t3 = time.perf_counter()
    it.film(out=host); t4 = time.perf_counter()
    print(k, "upload %.2f render %.2f train %.2f film %.2f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)), "samples", n, "cells", c)
