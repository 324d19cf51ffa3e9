"""One call of a hot-path entry point on a synthetic tensor, for ncu captures:
python tools/run_case.py {analyze|transition|spatial} [frames] [users] [tile counts, comma separated] [weighted 0/1] [repeats]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine

what = sys.argv[1] if len(sys.argv) > 1 else "analyze"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 450
U = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
tcs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [200]
use_w = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
dev = torch.device("cuda")
p = bench.synth_on_device(torch, F, U, 20265000, dev, chunk=32 if U > 200_000 else 256)
eng = get_engine(100, 200, tcs, EntropyConfig(fov_angle=90.0, power_factor=2.0, use_weight_distribution=use_w), dev)
for _ in range(reps):
    if what == "analyze":
        eng.analyze(p, want_per_k=False, want_assign0=True, want_pairs0=False)
    elif what == "transition":
        eng.transition(p, want_pairs0=False, want_per_k=False)
    else:
        eng.spatial(p, want_per_k=False)
torch.cuda.synchronize()
print("flags", eng.poll_flags())
