"""Weighted handle on a 1920x1080 video (2.08 M cells, 201 tiles, fov = 120): the global-table regime (per-cell weight
columns, round 2: allowed while the columns fit the budget) against the direct per-sample regime.
python tools/time_large_video.py [frames] [users]"""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from viewport_entropy_toolkit_b200 import Engine, EntropyConfig

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
U = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
p = bench.synth_on_device(torch, F, U, 31337, torch.device("cuda"))
for regime in ("auto", "direct"):
    t0 = time.perf_counter()
    e = Engine(1920, 1080, [200], EntropyConfig(fov_angle=120.0), regime=regime)
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    r = e.spatial(p)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = e.spatial(p)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"1920x1080, 201 tiles weighted, {F} frames x {U} users, regime={regime}: handle {t_create:.1f} s, spatial {ms:.2f} ms "
          f"= {F * U / ms / 1e6:.2f} G samples/s, entropy[0] {float(r.entropy[0]):.12f}, free memory {torch.cuda.mem_get_info()[0] / 2**30:.0f} GiB", flush=True)
    e.close()
