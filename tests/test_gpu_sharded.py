"""Multi-rank GPU test of the frame-sharded product path: `analyze_sharded` (Engine.analyze / vet_analyze on
every rank's frames + halo frame, one all-gather of [entropy | transition entropy | prev_count0] rows) must
return, bit for bit, what ONE rank returns for the whole video -- ragged shards, a rank without frames, missing
users, the halo frame, transition row offsets, and the precision mode of the weighted histogram pinned from the
global frame count.

Ranks run as separate processes.  With at least `world` GPUs on the box every rank takes its own GPU and the
collective is NCCL; on a one-GPU box the ranks share the GPU and the collective runs over gloo (NCCL refuses two
ranks on one device) -- the kernels, the sharding and the row packing are the same."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _video(F, U, seed, missing):
    rng = np.random.default_rng(seed)
    mu = np.clip(rng.normal(0.5, 0.15, U), 0, 1)[None] + np.cumsum(rng.normal(0, 0.02, (F, U)), 0)
    mv = np.clip(rng.normal(0.5, 0.10, U), 0, 1)[None] + np.cumsum(rng.normal(0, 0.012, (F, U)), 0)
    for a in (mu, mv):
        np.abs(a, out=a)
        a[a > 1] = 2 - a[a > 1]
        np.clip(a, 0, 1, out=a)
    p = np.stack([np.broadcast_to(np.arange(F)[:, None] * 0.1, (F, U)), mu, mv], -1).astype(np.float32)
    if missing:
        m = rng.uniform(size=(F, U)) < missing
        m[:, 0] = False
        p[m, 1] = np.nan
    return p


def _engine(cfg, device):
    from viewport_entropy_toolkit_b200 import Engine, EntropyConfig
    return Engine(100, 200, cfg["tcs"], EntropyConfig(fov_angle=cfg["fov"], use_weight_distribution=cfg["use_w"],
                                                      power_factor=2.0), device)


def _worker(rank, world, port, cfg, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    device = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(device)
    nccl = ndev >= world
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from viewport_entropy_toolkit_b200.distributed import analyze_sharded
    video = _video(cfg["F"], cfg["U"], cfg["seed"], cfg["missing"])
    eng = _engine(cfg, device)
    launches0 = eng.launch_count()

    def load(b, e):  # the rank loads its own frames, halo frame included
        return torch.from_numpy(np.ascontiguousarray(video[b:e])).to(device)

    res = analyze_sharded(eng, load, cfg["F"])
    flags = eng.poll_flags()
    np.savez(Path(out_dir) / f"r{rank}.npz", sp=res.sp_entropy.cpu().numpy(), tr=res.tr_entropy.cpu().numpy(),
             pc=res.prev_count0.cpu().numpy(), rng=[res.local_begin, res.local_end],
             hist0=res.hist0.cpu().numpy() if res.hist0 is not None else np.zeros((0, eng.num_tiles[0])),
             assign0=res.assign0.cpu().numpy() if res.assign0 is not None else np.zeros((0, cfg["U"]), np.uint16),
             flags=flags, launches=eng.launch_count() - launches0, nccl=int(nccl))
    dist.destroy_process_group()


CASES = [
    dict(world=2, F=9, U=3000, tcs=[200], fov=90.0, use_w=True, missing=0.0, seed=1),        # ragged 5 + 4, FP64 weighted histogram
    dict(world=3, F=7, U=2001, tcs=[20, 50], fov=120.0, use_w=False, missing=0.1, seed=2),   # missing users, two tile counts, odd U
    dict(world=4, F=3, U=500, tcs=[200], fov=90.0, use_w=True, missing=0.05, seed=3),        # rank 3 owns no frame
    dict(world=2, F=600, U=300, tcs=[200], fov=90.0, use_w=True, missing=0.02, seed=4),      # >= VET_I8_MIN_FRAMES: tensor-core mode pinned on both shards
]


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: f"w{c['world']}_F{c['F']}_U{c['U']}")
def test_sharded_analyze_equals_single_gpu(tmp_path, cfg):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.multiprocessing as mp
    world = cfg["world"]
    port = 29500 + (os.getpid() * 7 + world * 131 + cfg["F"]) % 2000
    mp.spawn(_worker, args=(world, port, cfg, str(tmp_path)), nprocs=world, join=True)

    dev = torch.device("cuda", 0)
    eng = _engine(cfg, dev)
    video = _video(cfg["F"], cfg["U"], cfg["seed"], cfg["missing"])
    sp, tr = eng.analyze(torch.from_numpy(video).to(dev), want_pairs0=False)
    assert eng.poll_flags() == 0
    sp_e, tr_e, pc = sp.entropy.cpu().numpy(), tr.entropy.cpu().numpy(), tr.prev_count0.cpu().numpy()
    hist0, assign0 = sp.hist0.cpu().numpy(), sp.assign0.cpu().numpy()
    covered = []
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert int(z["flags"]) == 0
        # the per-frame rows are complete and identical on every rank, bit for bit
        assert np.array_equal(z["sp"], sp_e, equal_nan=True)
        assert np.array_equal(z["tr"], tr_e, equal_nan=True)
        assert z["pc"].dtype == np.int32 and np.array_equal(z["pc"], pc)
        b, e = (int(v) for v in z["rng"])
        covered.append((b, e))
        # the per-user outputs stay with the owner of the frames
        assert np.array_equal(z["hist0"], hist0[b:e]) and np.array_equal(z["assign0"], assign0[b:e])
        assert (int(z["launches"]) > 0) == (e > b), "a rank with frames must have launched the CUDA kernels"
    assert covered[0][0] == 0 and covered[-1][1] == cfg["F"] and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    eng.close()
