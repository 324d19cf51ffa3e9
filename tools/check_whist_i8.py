"""GPU check of the tensor-core weighted histogram (k_whist_i8) against the FP64 kernel (k_whist):
same engine, same input, VET_WHIST_IMPL switched between calls.  Prints max differences and timings.

    python tools/check_whist_i8.py [small|mid|big|hot|huge ...]
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine

dev = torch.device("cuda")


def run(eng, p, impl):
    os.environ["VET_WHIST_IMPL"] = impl
    r = eng.spatial(p)
    torch.cuda.synchronize()
    return r


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


def case(name, F, U, tcs, fov=90.0, pf=2.0, hot=0, timing=False):
    p = bench.synth_on_device(torch, F, U, 7, dev)
    if hot:  # pile `hot` users of every frame onto one cell: counts >= 256 (plane 1) or >= 65536 (FP64 fallback)
        p[:, :hot, 1] = 0.5
        p[:, :hot, 2] = 0.5
    eng = get_engine(100, 200, tcs, EntropyConfig(fov, True, pf), dev)
    a = run(eng, p, "fp64")
    b = run(eng, p, "i8")
    b2 = run(eng, p, "i8")
    h_a, h_b = a.hist0, b.hist0
    den = h_a.abs().clamp_min(1e-300)
    rel = ((h_a - h_b).abs() / den)
    mask = h_a.abs() > 1e-6
    out = dict(case=name, F=F, U=U, tiles=tcs,
               hist_max_abs=float((h_a - h_b).abs().max()),
               hist_max_rel_above_1e6=float(rel[mask].max()) if mask.any() else 0.0,
               entropy_max_rel=float(((a.entropy - b.entropy).abs() / a.entropy.abs().clamp_min(1e-300)).max()),
               per_k_max_abs=float((a.per_k - b.per_k).abs().max()),
               rerun_identical=bool(torch.equal(b.hist0, b2.hist0) and torch.equal(b.entropy, b2.entropy)),
               assign_equal=bool(torch.equal(a.assign0, b.assign0)))
    if timing:
        for impl in ("fp64", "i8"):
            os.environ["VET_WHIST_IMPL"] = impl
            eng.profile(True)
            ms = timeit(lambda: eng.spatial(p, want_per_k=False))
            out[f"ms_{impl}"] = ms
            out[f"prof_{impl}"] = eng.profile_read()
            eng.profile(False)
    print(json.dumps(out), flush=True)
    del p


def dirty_case():
    """planes 1/2 written by a 'hot' call must be cleaned by the next call on the same rows"""
    eng = get_engine(100, 200, [200], EntropyConfig(90.0, True, 2.0), dev)
    hot = bench.synth_on_device(torch, 260, 70000, 11, dev)
    hot[:, :66000, 1] = 0.25
    hot[:, :66000, 2] = 0.75
    cold = bench.synth_on_device(torch, 200, 70000, 12, dev)
    run(eng, hot, "i8")
    b = run(eng, cold, "i8")
    a = run(eng, cold, "fp64")
    print(json.dumps(dict(case="dirty_rows", hist_max_abs=float((a.hist0 - b.hist0).abs().max()),
                          entropy_max_rel=float(((a.entropy - b.entropy).abs() / a.entropy.abs()).max()))), flush=True)


which = sys.argv[1:] or ["small", "mid"]
if "small" in which:
    case("small", 300, 5000, [200])
    case("small_k2", 130, 3000, [20, 200], fov=120.0)
if "mid" in which:
    case("mid_plane1", 260, 20000, [200], hot=3000)
    case("mid_fallback", 130, 70000, [200], hot=66000)
    case("mid_T1001", 200, 4000, [1000], fov=60.0)
    case("mid_chunks", 130, 300000, [200], hot=70000)
    dirty_case()
if "big" in which:
    case("c3", 3600, 100000, [200], timing=True)
