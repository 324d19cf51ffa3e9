"""configs[1] (10k users x 1800 frames, tile counts 20/50/100/200, unweighted): Engine.spatial on a side stream (CUDA-graph
replays), ms per call.  python tools/time_c2.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine

dev = torch.device("cuda")
torch.cuda.set_stream(torch.cuda.Stream(dev))
wl = bench.WORKLOADS["c2"]
p = bench.synth_on_device(torch, wl["F"], wl["U"], 20265000, dev)
eng = get_engine(100, 200, wl["tile_counts"], EntropyConfig(fov_angle=wl["fov"], use_weight_distribution=wl["use_w"], power_factor=wl["pf"]), dev)
for graph in ("on", "off"):
    eng.set_option("cuda_graph", graph)
    for _ in range(6):
        r = eng.spatial(p)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        r = eng.spatial(p)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    print(f"cuda_graph={graph}: {ms:.4f} ms per call, {wl['F'] * wl['U'] * 14 / ms / 1e6 / 6535.7:.3f} of the roofline, "
          f"graph replays {eng.graph_replays()}, flags {eng.poll_flags()}, entropy[0:2] {r.entropy[:2].tolist()}", flush=True)
