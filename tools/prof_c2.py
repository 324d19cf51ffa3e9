"""configs[1] (10k users x 1800 frames, tile_counts=[20,50,100,200], unweighted) three times, for ncu:
ncu --set full -k regex:k_stream_tiles -s 2 -c 1 python tools/prof_c2.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine

wl = bench.WORKLOADS["c2"]
p = bench.synth_on_device(torch, wl["F"], wl["U"], 1, torch.device("cuda"))
eng = get_engine(100, 200, wl["tile_counts"], EntropyConfig(wl["fov"], wl["use_w"], wl["pf"]), torch.device("cuda"))
for _ in range(3):
    eng.spatial(p)
torch.cuda.synchronize()
print("flags", eng.poll_flags())
