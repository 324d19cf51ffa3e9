"""World-size-2 (3, 4) gloo tests of the frame-sharding logic on CPU: shard bounds, the halo
frame, the single all-gather with ragged and empty shards, the row packing of
`analyze_sharded`.  The per-rank compute is the CPU oracle here (no GPU in this
container), behind a stand-in with Engine's interface; the same `analyze_sharded` around
the real Engine.analyze (vet_analyze) is what `bench.py` times at every N
(`extra.c5_sharded`) and what tests/test_gpu_sharded.py compares with the single-GPU run
on the GPU box."""
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


class _OracleEngine:
    """Engine's interface (analyze, options, num_tiles, device) over the CPU oracle."""
    device = torch.device("cpu")

    def __init__(self, tile_counts, fov, use_w):
        from oracle import vet_oracle as orc
        self.orc, self.tcs, self.fov, self.use_w = orc, tile_counts, fov, use_w
        self.num_tiles = [len(orc.lattice(n)) for n in tile_counts]
        self.pinned = []

    def get_option(self, name):
        return "auto"

    def set_option(self, name, value):
        self.pinned.append((name, value))

    def analyze(self, packed, mode="literal", **kw):
        from types import SimpleNamespace as NS
        p = packed.numpy()
        s = self.orc.spatial_analyzer(p, 100, 200, self.tcs, self.fov, self.use_w, 2.0)
        sp = NS(entropy=torch.from_numpy(s["entropy"]), hist0=torch.from_numpy(s["hist0"]),
                assign0=torch.from_numpy(s["assign0"].astype(np.int32)))
        if p.shape[0] > 1:
            t = self.orc.transition_analyzer(p, 100, 200, self.tcs)
            tr = NS(entropy=torch.from_numpy(t["entropy"]), prev_count0=torch.from_numpy(t["prev_count0"]))
        else:
            tr = NS(entropy=torch.empty(0, dtype=torch.float64), prev_count0=torch.empty((0, self.num_tiles[0]), dtype=torch.int32))
        return sp, tr


def _analyze_worker(rank, world, port, F, U, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from viewport_entropy_toolkit_b200.distributed import ShardedAnalyzer, analyze_read_range, analyze_sharded
    rng = np.random.default_rng(11)
    packed = np.stack([np.zeros((F, U)), rng.uniform(0, 1, (F, U)), rng.uniform(0, 1, (F, U))], -1).astype(np.float32)
    eng = _OracleEngine([20], 120.0, True)
    res = analyze_sharded(eng, lambda b, e: torch.from_numpy(packed[b:e]), F)
    # the two-phase form used by bench.py (kernels of the next call may overlap the collective)
    sa = ShardedAnalyzer(eng, F)
    rb, b, e = analyze_read_range(F, rank, world)
    sa.start(torch.from_numpy(packed[rb:e]))
    res2 = sa.finish()
    assert torch.equal(res.sp_entropy, res2.sp_entropy) and torch.equal(res.prev_count0, res2.prev_count0)
    np.savez(Path(out_dir) / f"a{rank}.npz", sp=res.sp_entropy.numpy(), tr=res.tr_entropy.numpy(), pc=res.prev_count0.numpy(),
             rng=[res.local_begin, res.local_end], hist0=res.hist0.numpy() if res.hist0 is not None else np.zeros((0, 21)),
             pinned=len(eng.pinned))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,F", [(2, 7), (3, 5), (4, 3)])
def test_analyze_sharded_rows(tmp_path, world, F):
    """analyze_sharded: every rank ends with the complete spatial rows [F], transition rows [F-1] and int32
    prev_count0 [F-1, T0]; row offsets across the halo frame; a rank without frames (world 4, 3 frames)."""
    from oracle import vet_oracle as orc
    U = 30
    port = 31500 + (os.getpid() + world * 11 + F) % 2000
    mp.spawn(_analyze_worker, args=(world, port, F, U, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(11)
    packed = np.stack([np.zeros((F, U)), rng.uniform(0, 1, (F, U)), rng.uniform(0, 1, (F, U))], -1).astype(np.float32)
    ref_s = orc.spatial_analyzer(packed, 100, 200, [20], 120.0, True, 2.0)
    ref_t = orc.transition_analyzer(packed, 100, 200, [20])
    for r in range(world):
        z = np.load(tmp_path / f"a{r}.npz")
        assert np.array_equal(z["sp"], ref_s["entropy"])
        assert np.array_equal(z["tr"], ref_t["entropy"], equal_nan=True)
        assert z["pc"].dtype == np.int32 and np.array_equal(z["pc"], ref_t["prev_count0"])
        b, e = (int(v) for v in z["rng"])
        assert np.array_equal(z["hist0"], ref_s["hist0"][b:e])
        assert int(z["pinned"]) == (4 if e > b else 0)   # weighted-kernel mode pinned and released around every analyze


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
    from viewport_entropy_toolkit_b200.distributed import analyze_read_range
    assert [analyze_read_range(3600, r, 8) for r in range(8)][:2] == [(0, 0, 450), (449, 450, 900)]   # 450 frames + 1 halo
    assert [analyze_read_range(3, r, 4) for r in range(4)] == [(0, 0, 1), (0, 1, 2), (1, 2, 3), (3, 3, 3)]
