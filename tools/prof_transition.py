"""Small transition-only run for ncu: python tools/prof_transition.py 200|500|1000 [frames]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine

tc = [int(x) for x in sys.argv[1].split(",")]
F = int(sys.argv[2]) if len(sys.argv) > 2 else 297
dev = torch.device("cuda")
p = bench.synth_on_device(torch, F, 100_000, 4, dev)
eng = get_engine(100, 200, tc, EntropyConfig(use_weight_distribution=False), dev)
for _ in range(2):
    r = eng.transition(p, want_pairs0=False, want_per_k=False)
torch.cuda.synchronize()
print(float(r.entropy.sum()))
