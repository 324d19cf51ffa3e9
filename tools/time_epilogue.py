"""Where the weighted epilogue of vet_analyze goes on the configs[4] shard (1M users x 450 frames, 200 tiles):
FP64 kernel against the tensor-core kernel (split over the cells), beside the transition kernels (side stream) or after
them, and the spatial stage alone.  python tools/time_epilogue.py [frames] [users]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine

F = int(sys.argv[1]) if len(sys.argv) > 1 else 450
U = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dev = torch.device("cuda")
torch.cuda.set_stream(torch.cuda.Stream(dev))
p = bench.synth_on_device(torch, F, U, 20265000, dev, chunk=32 if U > 200_000 else 256)
eng = get_engine(100, 200, [200], EntropyConfig(fov_angle=90.0, power_factor=2.0), dev)


def timed(fn, n=8):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for wk in ("fp64", "i8", "auto"):
    eng.set_option("weighted_kernel", wk)
    row = {}
    for ov in ("on", "off"):
        eng.set_option("analyze_overlap", ov)
        row[f"analyze overlap={ov}"] = timed(lambda: eng.analyze(p, want_per_k=False, want_assign0=True, want_pairs0=False))
    eng.set_option("analyze_overlap", "on")
    row["spatial"] = timed(lambda: eng.spatial(p, want_per_k=False, want_hist0=False))
    row["transition"] = timed(lambda: eng.transition(p, want_pairs0=False, want_per_k=False))
    print(f"F={F} U={U} weighted_kernel={wk}: " + ", ".join(f"{k} {v:.4f} ms" for k, v in row.items()), flush=True)
