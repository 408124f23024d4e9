"""One steady-state guided training step of a bench workload between cudaProfilerStart / Stop, one wavefront lane and no stream
overlap, for `ncu --profile-from-start off` (launch list or --set full of exactly that step).
usage: profile_step.py [workload] [pretrain] [spp] [max_cell_samples]   (workload as bench.py --workload)"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench
import torch
name = sys.argv[1] if len(sys.argv) > 1 else "cornell_caustic_1024"
pretrain = int(sys.argv[2]) if len(sys.argv) > 2 else 8
args = bench.parse_args([])
args.workload = name
args.spp_per_step = int(sys.argv[3]) if len(sys.argv) > 3 else (1 if name == "mesh_10m" else 4)
args.guided_distance = name == "medium_1024"
if len(sys.argv) > 4:
    args.max_cell_samples = int(sys.argv[4])  # a smaller split threshold grows the larger field of a multi-GPU job on one GPU
pkg, sb, desc = bench.workload(name)
from b200pg import api
p = bench.guided_params(pkg, args)
it = api.Integrator(api.Scene.from_builder(sb), p, device=0)
it.set_option("lanes", 1)
it.set_option("overlap_shadow", 0)
spp = args.spp_per_step


def step(k):
    it.guiding_mode(True, k > 0)
    it.progression(k * spp, spp)
    return it.train_fused(args.em_iters)


for k in range(pretrain):
    step(k)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
n, c = step(pretrain)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(desc, "| samples", n, "cells", c, it.stage_times())
