"""World-size-2 (and 3) gloo tests of the frame-sharding logic on CPU: shard bounds,
the transition halo frame, the single all-gather with ragged shards.  The per-rank
compute is the CPU oracle here (no GPU in this container); on the GPU box the same
`run_sharded` wraps Engine.spatial / Engine.transition (bench.py, N > 1)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _worker(rank, world, port, F, U, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import vet_oracle as orc
    from viewport_entropy_toolkit_b200.distributed import run_sharded
    rng = np.random.default_rng(5)
    packed = np.stack([np.zeros((F, U)), rng.uniform(0, 1, (F, U)), rng.uniform(0, 1, (F, U))], -1).astype(np.float32)

    def spatial_local(b, e):
        r = orc.spatial_analyzer(packed[b:e], 100, 200, [20], 120.0, False, 2.0)
        return torch.from_numpy(r["entropy"]), torch.from_numpy(r["hist0"])

    def transition_local(b, e):
        if e - b < 2:
            return torch.empty(0, dtype=torch.float64), torch.empty((0, 21), dtype=torch.float64)
        r = orc.transition_analyzer(packed[b:e], 100, 200, [20])
        return torch.from_numpy(r["entropy"]), torch.from_numpy(r["prev_count0"].astype(np.float64))

    sp = run_sharded(F, spatial_local)
    tr = run_sharded(F, transition_local, transition=True)
    np.savez(Path(out_dir) / f"r{rank}.npz", sp_e=sp.entropy.numpy(), sp_h=sp.rows.numpy(), tr_e=tr.entropy.numpy(),
             tr_c=tr.rows.numpy(), sp_range=[sp.local_begin, sp.local_end], tr_range=[tr.local_begin, tr.local_end])
    dist.destroy_process_group()


@pytest.mark.parametrize("world,F", [(2, 7), (2, 8), (3, 5)])
def test_sharded_equals_single_process(tmp_path, world, F):
    from oracle import vet_oracle as orc
    U = 40
    port = 29500 + (os.getpid() + world * 7 + F) % 2000
    mp.spawn(_worker, args=(world, port, F, U, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    packed = np.stack([np.zeros((F, U)), rng.uniform(0, 1, (F, U)), rng.uniform(0, 1, (F, U))], -1).astype(np.float32)
    ref_s = orc.spatial_analyzer(packed, 100, 200, [20], 120.0, False, 2.0)
    ref_t = orc.transition_analyzer(packed, 100, 200, [20])
    covered = []
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["sp_e"], ref_s["entropy"]) and np.array_equal(z["sp_h"], ref_s["hist0"])
        assert np.array_equal(z["tr_e"], ref_t["entropy"], equal_nan=True)
        assert np.array_equal(z["tr_c"], ref_t["prev_count0"])
        covered.append(tuple(z["sp_range"]))
    assert covered[0][0] == 0 and covered[-1][1] == F and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_ranges():
    from viewport_entropy_toolkit_b200.distributed import frame_range, transition_range
    assert [frame_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [frame_range(3600, r, 8) for r in range(8)][-1] == (3150, 3600)
    assert [transition_range(10, r, 2) for r in range(2)] == [(0, 5, 6), (5, 9, 10)]   # halo frame included
    assert transition_range(1, 0, 2) == (0, 0, 0)
    assert [frame_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
