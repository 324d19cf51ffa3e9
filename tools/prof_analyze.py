"""vet_analyze on the configs[4] shard (1M users x 450 frames, 200 tiles, weighted) for ncu:
python tools/prof_analyze.py [frames] [users] [iterations]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine

F = int(sys.argv[1]) if len(sys.argv) > 1 else 450
U = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda")
p = bench.synth_on_device(torch, F, U, 20265000, dev, chunk=32 if U > 200_000 else 256)
eng = get_engine(100, 200, [200], EntropyConfig(fov_angle=90.0, power_factor=2.0), dev)
for _ in range(n):
    sp, tr = eng.analyze(p, want_per_k=False, want_assign0=True, want_pairs0=False)
torch.cuda.synchronize()
print(float(sp.entropy.sum()), float(tr.entropy.sum()))
