"""Ad-hoc device timings of the BASELINE configs other than the headline one."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


dev = torch.device("cuda")
out = {}
which = sys.argv[1:] or ["c2", "c5shard", "tr"]
if "c2" in which:
    wl = bench.WORKLOADS["c2"]
    p = bench.synth_on_device(torch, wl["F"], wl["U"], 1, dev)
    eng = get_engine(100, 200, wl["tile_counts"], EntropyConfig(wl["fov"], wl["use_w"], wl["pf"]), dev)
    eng.profile(True)
    ms = timeit(lambda: eng.spatial(p))
    out["c2_spatial"] = dict(ms=ms, gsamples=wl["F"] * wl["U"] / ms / 1e6, prof=eng.profile_read())
    eng.profile(False)
    del p
if "c5shard" in which:
    wl = bench.WORKLOADS["c5shard"]
    p = bench.synth_on_device(torch, wl["F"], wl["U"], 2, dev, chunk=32)
    eng = get_engine(100, 200, wl["tile_counts"], EntropyConfig(wl["fov"], wl["use_w"], wl["pf"]), dev)
    eng.profile(True)
    ms = timeit(lambda: eng.spatial(p, want_per_k=False), n=3, warm=1)
    out["c5shard_spatial"] = dict(ms=ms, gsamples=wl["F"] * wl["U"] / ms / 1e6, prof=eng.profile_read())
    eng.profile(False)
    del p
if "tr" in which:
    for U, F, tcs in [(100_000, 33, [200]), (100_000, 33, [200, 500, 1000]), (1_000_000, 9, [200])]:
        p = bench.synth_on_device(torch, F, U, 3, dev, chunk=16)
        eng = get_engine(100, 200, tcs, EntropyConfig(use_weight_distribution=False), dev)
        eng.profile(True)
        ms = timeit(lambda: eng.transition(p, want_pairs0=False, want_per_k=False), n=2, warm=1)
        out[f"transition_U{U}_F{F}_{tcs}"] = dict(ms=ms, gpairs=(F - 1) * U / ms / 1e6, prof=eng.profile_read())
        eng.profile(False)
        del p
if "c4" in which:  # BASELINE configs[3] at full size on one GPU, and its single-tile-count pieces
    p = bench.synth_on_device(torch, 3600, 100_000, 4, dev)
    for tcs in ([200, 500, 1000], [200], [500], [1000]):
        eng = get_engine(100, 200, tcs, EntropyConfig(use_weight_distribution=False), dev)
        eng.profile(True)
        ms = timeit(lambda: eng.transition(p, want_pairs0=False, want_per_k=False), n=2, warm=1)
        out[f"c4_transition_{tcs}"] = dict(ms=ms, gpairs=3599 * 100_000 / ms / 1e6, prof=eng.profile_read())
        eng.profile(False)
    del p
if "mid" in which:  # the reference README's custom configuration: 200x400 video, tile_counts=[50,100,200], default weighting
    F, U = 900, 100_000
    p = bench.synth_on_device(torch, F, U, 6, dev)
    for regime in ("global", "direct"):
        if regime == "direct":
            p = p[:30].contiguous()   # the per-sample regime is O(U*T) per frame: time a slice
        from viewport_entropy_toolkit_b200.engine import Engine
        eng = Engine(200, 400, [50, 100, 200], EntropyConfig(), dev, regime="direct" if regime == "direct" else "auto")
        eng.profile(True)
        ms = timeit(lambda: eng.spatial(p, want_per_k=False), n=2, warm=1)
        out[f"mid_200x400_{regime}"] = dict(frames=int(p.shape[0]), users=U, ms=ms, gsamples=p.shape[0] * U / ms / 1e6,
                                            prof=eng.profile_read())
        eng.close()
    del p
if "midu" in which:  # the same video, unweighted
    from viewport_entropy_toolkit_b200.engine import Engine
    F, U = 900, 100_000
    p = bench.synth_on_device(torch, F, U, 6, dev)
    for dims in ((200, 400), (1920, 1080)):
        eng = Engine(dims[0], dims[1], [50, 100, 200], EntropyConfig(use_weight_distribution=False), dev)
        eng.profile(True)
        ms = timeit(lambda: eng.spatial(p, want_per_k=False), n=2, warm=1)
        out[f"mid_{dims[0]}x{dims[1]}_unweighted"] = dict(frames=F, users=U, ms=ms, gsamples=F * U / ms / 1e6, prof=eng.profile_read())
        eng.close()
    del p
print(json.dumps(out))
