#!/usr/bin/env python
"""Generates tests/golden/*.npz from the LIVE reference (build container only).

Run:  python tests/golden/make_golden.py            (needs /root/reference)

The reference cannot travel to the GPU box, so its outputs on seeded inputs are
committed here as small fixtures.  Every array below is produced by calling the
reference's own functions (through oracle/_refshim.py, which only stubs the
absent pyvista/matplotlib imports); nothing from oracle/vet_oracle.py is used
to produce them.
"""
import os
import sys
import tempfile
import warnings
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent

from oracle._refshim import load_reference  # noqa: E402

TILE_COUNTS_ALL = [20, 50, 100, 200, 250, 500, 1000]
W0, H0 = 100, 200


def _ref():
    load_reference()
    import viewport_entropy_toolkit as vet
    from viewport_entropy_toolkit import utilities as U
    return vet, U


def ref_cell_vector(px, py, W, H):
    """The reference's own per-row decode chain for one (px,py):
    pixel_to_spherical (DU:264-286) -> round/wrap (DU:390-397) ->
    RadialPoint -> Vector.from_spherical (DU:399-403)."""
    vet, U = _ref()
    rp = U.pixel_to_spherical(vet.Point(np.float64(px), np.float64(py)), W, H)
    lon = round(float(rp.lon), 1)
    lat = round(float(rp.lat), 1)
    if lon <= -180:
        lon = (lon + 360) % 360 - 180
    if lat <= -90:
        lat = (lat + 180) % 180 - 90
    vet.RadialPoint(lon=lon, lat=lat)
    v = vet.Vector.from_spherical(lon, lat)
    return lon, lat, (v.x, v.y, v.z)


def gen_lattices():
    vet, U = _ref()
    out = {}
    for n in [1, 2, 3, 7, 20, 21, 50, 100, 200, 250, 500, 1000]:
        L = U.generate_fibonacci_lattice(n)
        out[f"n{n}"] = np.array([[v.x, v.y, v.z] for v in L], dtype=np.float64)
    np.savez_compressed(OUT / "lattices.npz", **out)
    print("lattices", {k: v.shape for k, v in out.items()})


def gen_decode():
    out = {}
    for (W, H) in [(W0, H0), (200, 400), (64, 32)]:
        lon = np.empty(W + 1)
        lat = np.empty(H + 1)
        for px in range(W + 1):
            lon[px] = ref_cell_vector(px, 0, W, H)[0]
        for py in range(H + 1):
            lat[py] = ref_cell_vector(0, py, W, H)[1]
        out[f"lon_{W}x{H}"] = lon
        out[f"lat_{W}x{H}"] = lat
    # full cell-vector grid at the default dims, and at 200x400 (README example)
    for (W, H) in [(W0, H0), (200, 400)]:
        cv = np.empty((H + 1, W + 1, 3))
        for py in range(H + 1):
            for px in range(W + 1):
                cv[py, px] = ref_cell_vector(px, py, W, H)[2]
        out[f"cellvec_{W}x{H}"] = cv
    # axis tables only for video-like dims (grid too large to store)
    for (W, H) in [(1920, 1080), (3840, 1920)]:
        out[f"lon_{W}x{H}"] = np.array([ref_cell_vector(px, 0, W, H)[0] for px in range(W + 1)])
        out[f"lat_{W}x{H}"] = np.array([ref_cell_vector(0, py, W, H)[1] for py in range(H + 1)])
    # normalize_to_pixel known answers incl. the 0.29*100 -> 28 quirk
    vet, U = _ref()
    mus = np.array([0.0, 1.0, 0.5, 0.29, 0.57, 0.58, 0.35, 0.07, 0.999, 0.001, 0.123456, 0.654321, 0.75, 0.25])
    out["ntp_in"] = mus
    out["ntp_100"] = U.normalize_to_pixel(mus, 100)
    out["ntp_200"] = U.normalize_to_pixel(mus, 200)
    out["ntp_1920"] = U.normalize_to_pixel(mus, 1920)
    mus32 = mus.astype(np.float32)
    out["ntp32_100"] = U.normalize_to_pixel(mus32.astype(np.float64), 100)
    out["ntp32_200"] = U.normalize_to_pixel(mus32.astype(np.float64), 200)
    np.savez_compressed(OUT / "decode.npz", **out)
    print("decode", sorted(out))


def _nearest_chunk(args):
    n, vecs = args
    vet, U = _ref()
    L = U.generate_fibonacci_lattice(n)
    return np.array([U.find_nearest_tile(vet.Vector(*v), L) for v in vecs], dtype=np.uint16)


def gen_nearest():
    """Nearest tile (reference find_nearest_tile, EU:89-106) of EVERY default-grid
    cell for all seven tile counts: the exhaustive domain."""
    cv = np.load(OUT / "decode.npz")[f"cellvec_{W0}x{H0}"].reshape(-1, 3)
    out = {}
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for n in TILE_COUNTS_ALL:
            chunks = np.array_split(cv, 64)
            res = list(ex.map(_nearest_chunk, [(n, c) for c in chunks]))
            out[f"lut_n{n}"] = np.concatenate(res)
            print("nearest", n, out[f"lut_n{n}"][:8], flush=True)
    # a sample of the 200x400 grid for n=20,200
    cv2 = np.load(OUT / "decode.npz")["cellvec_200x400"].reshape(-1, 3)
    rng = np.random.default_rng(7)
    sel = rng.choice(len(cv2), 4000, replace=False)
    out["sel_200x400"] = sel
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for n in [20, 200]:
            chunks = np.array_split(cv2[sel], 32)
            out[f"lut200x400_n{n}"] = np.concatenate(list(ex.map(_nearest_chunk, [(n, c) for c in chunks])))
    # arbitrary (non-grid, non-unit) vectors
    arb = rng.normal(size=(512, 3)) * rng.uniform(0.1, 5.0, size=(512, 1))
    out["arb_vecs"] = arb
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for n in [20, 200, 1000]:
            chunks = np.array_split(arb, 16)
            out[f"arb_n{n}"] = np.concatenate(list(ex.map(_nearest_chunk, [(n, c) for c in chunks])))
    np.savez_compressed(OUT / "nearest.npz", **out)


def _weights_chunk(args):
    n, vecs, fov, pf = args
    vet, U = _ref()
    L = U.generate_fibonacci_lattice(n)
    idx = {c: i for i, c in enumerate(L)}
    cfg = U.EntropyConfig(fov_angle=fov, use_weight_distribution=True, power_factor=pf)
    out = np.zeros((len(vecs), len(L)))
    for r, v in enumerate(vecs):
        for c, w in U.calculate_tile_weights(vet.Vector(*v), L, cfg).items():
            out[r, idx[c]] = w
    return out


def gen_weights():
    """calculate_tile_weights (EU:108-144) rows for a sample of default-grid cells."""
    cv = np.load(OUT / "decode.npz")[f"cellvec_{W0}x{H0}"].reshape(-1, 3)
    rng = np.random.default_rng(11)
    sel = np.sort(rng.choice(len(cv), 160, replace=False))
    sel[:6] = [0, 100, 50 + 100 * 101, 100 + 100 * 101, 100 + 200 * 101, 20300]
    sel = np.unique(sel)
    out = {"sel": sel}
    jobs = [(20, 120.0, 2.0), (200, 90.0, 2.0), (200, 120.0, 2.0), (50, 60.0, 1.5), (200, 360.0, 3.0), (1000, 120.0, 2.0)]
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for (n, fov, pf) in jobs:
            chunks = np.array_split(cv[sel], 16)
            res = list(ex.map(_weights_chunk, [(n, c, fov, pf) for c in chunks]))
            out[f"w_n{n}_fov{int(fov)}_pf{pf}"] = np.concatenate(res)
            print("weights", n, fov, pf, flush=True)
    np.savez_compressed(OUT / "weights.npz", **out)


def synth_packed(F, U, seed, missing=0.0, iid=False):
    """Seeded synthetic (time, 2dmu, 2dmv) float32 [F,U,3] (SURVEY 8d generator,
    numpy version): gaussian start + reflected random walk, or iid uniform."""
    rng = np.random.default_rng(seed)
    if iid:
        mu = rng.uniform(0, 1, size=(F, U))
        mv = rng.uniform(0, 1, size=(F, U))
    else:
        mu = np.clip(rng.normal(0.5, 0.15, size=U), 0, 1)[None, :] + np.cumsum(rng.normal(0, 0.010, size=(F, U)), axis=0)
        mv = np.clip(rng.normal(0.5, 0.10, size=U), 0, 1)[None, :] + np.cumsum(rng.normal(0, 0.006, size=(F, U)), axis=0)
        mu = np.abs(mu); mu = np.where(mu > 1, 2 - mu, mu)
        mv = np.abs(mv); mv = np.where(mv > 1, 2 - mv, mv)
        mu = np.clip(mu, 0, 1); mv = np.clip(mv, 0, 1)
    t = np.broadcast_to((np.arange(F) * 0.1)[:, None], (F, U))
    p = np.stack([t, mu, mv], axis=-1).astype(np.float32)
    edge = [(0.5, 0.5), (0.0, 0.5), (1.0, 0.5), (1.0, 1.0), (0.29, 0.57), (0.999, 0.001), (0.123456, 0.654321), (0.75, 0.25)]
    for u, (a, b) in enumerate(edge[:U]):
        p[0, u, 1] = a
        p[0, u, 2] = b
    if missing > 0:
        m = rng.uniform(size=(F, U)) < missing
        m[:, 0] = False  # keep user 0 always present so no frame is empty
        p[m, 1] = np.nan
        p[m, 2] = np.nan
    return p


def ref_vectors_from_packed(packed, W, H):
    """packed[F,U,3] -> list over frames of {user_name: Vector} via the reference decode chain."""
    vet, U = _ref()
    F, Un, _ = packed.shape
    frames = []
    for f in range(F):
        d = {}
        mu = packed[f, :, 1].astype(np.float64)
        mv = packed[f, :, 2].astype(np.float64)
        ok = ~(np.isnan(mu) | np.isnan(mv))
        px = U.normalize_to_pixel(np.where(ok, mu, 0.0), W)
        py = U.normalize_to_pixel(np.where(ok, mv, 0.0), H)
        for u in range(Un):
            if ok[u]:
                d[f"u{u:05d}"] = vet.Vector(*ref_cell_vector(px[u], py[u], W, H)[2])
        frames.append(d)
    return frames


def _spatial_frame(args):
    d_items, tile_counts, fov, use_w, pf = args
    vet, U = _ref()
    d = {k: vet.Vector(*v) for k, v in d_items}
    cfg = U.EntropyConfig(fov_angle=fov, use_weight_distribution=use_w, power_factor=pf)
    res = []
    for n in tile_counts:
        L = U.generate_fibonacci_lattice(n)
        idx = {c: i for i, c in enumerate(L)}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            e, wts, asg = U.compute_spatial_entropy(d, L, cfg)
        hist = np.zeros(len(L))
        for c, w in wts.items():
            hist[idx[c]] = w
        res.append((float(e), hist, [asg[k] for k, _ in d_items]))
    return res


def _transition_pair(args):
    prev_items, cur_items, tile_counts = args
    vet, U = _ref()
    prior = {k: vet.Vector(*v) for k, v in prev_items}
    cur = {k: vet.Vector(*v) for k, v in cur_items}
    cfg = U.EntropyConfig()
    res = []
    for n in tile_counts:
        L = U.generate_fibonacci_lattice(n)
        idx = {c: i for i, c in enumerate(L)}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            e, wts, asg = U.compute_transition_entropy(prior, cur, L, cfg, 120)
        m = np.zeros(len(L), dtype=np.int32)
        for c, w in wts.items():
            m[idx[c]] = w
        res.append((float(e), m, asg))
    return res


def gen_frames():
    """compute_spatial_entropy / compute_transition_entropy (EU:147-332) on
    seeded packed tensors, through the reference decode chain."""
    cases = [
        # name, F, U, seed, missing, iid, tile_counts, fov, use_w, pf
        ("c_small_w120", 6, 24, 101, 0.0, False, [20, 50], 120.0, True, 2.0),
        ("c_small_unw", 6, 24, 102, 0.0, False, [20, 50, 100, 200], 120.0, False, 2.0),
        ("c_w90_t200", 4, 40, 103, 0.0, False, [200], 90.0, True, 2.0),
        ("c_missing", 8, 16, 104, 0.3, False, [20, 200], 120.0, True, 2.0),
        ("c_iid_unw", 5, 300, 105, 0.0, True, [20, 50], 120.0, False, 2.0),
        ("c_iid_w", 3, 64, 106, 0.1, True, [50, 250], 100.0, True, 1.5),
        ("c_oneuser", 3, 1, 107, 0.0, False, [20], 120.0, False, 2.0),
        ("c_t1000", 3, 12, 108, 0.0, False, [1000], 120.0, True, 2.0),
    ]
    out = {}
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for (name, F, Un, seed, miss, iid, tcs, fov, use_w, pf) in cases:
            packed = synth_packed(F, Un, seed, miss, iid)
            frames = ref_vectors_from_packed(packed, W0, H0)
            items = [[(k, (v.x, v.y, v.z)) for k, v in d.items()] for d in frames]
            res = list(ex.map(_spatial_frame, [(it, tcs, fov, use_w, pf) for it in items]))
            K = len(tcs)
            per_k = np.array([[res[f][k][0] for f in range(F)] for k in range(K)])
            out[f"{name}/packed"] = packed
            out[f"{name}/tile_counts"] = np.array(tcs)
            out[f"{name}/cfg"] = np.array([fov, float(use_w), pf])
            out[f"{name}/sp_per_k"] = per_k
            # SA:156 average exactly as the reference accumulates it
            ent = []
            for f in range(F):
                tot = 0
                for k in range(K):
                    tot += res[f][k][0]
                ent.append(tot / K)
            out[f"{name}/sp_entropy"] = np.array(ent)
            out[f"{name}/sp_hist0"] = np.array([res[f][0][1] for f in range(F)])
            asg = np.full((F, Un), 0xFFFF, dtype=np.uint16)
            for f in range(F):
                users = [int(k[1:]) for k, _ in items[f]]
                asg[f, users] = res[f][0][2]
            out[f"{name}/sp_assign0"] = asg
            # transitions (skip when a pair has no common user -> ZeroDivisionError in the reference)
            tres = list(ex.map(_transition_pair, [(items[f - 1], items[f], tcs) for f in range(1, F)]))
            tper_k = np.array([[tres[r][k][0] for r in range(F - 1)] for k in range(K)])
            out[f"{name}/tr_per_k"] = tper_k
            tent = []
            for r in range(F - 1):
                tot = 0
                for k in range(K):
                    tot += tres[r][k][0]
                tent.append(tot / K)
            out[f"{name}/tr_entropy"] = np.array(tent)
            out[f"{name}/tr_prev_count0"] = np.array([tres[r][0][1] for r in range(F - 1)])
            pairs = np.full((F - 1, Un, 2), 0xFFFF, dtype=np.uint16)
            for r in range(F - 1):
                for k, pc in tres[r][0][2].items():
                    pairs[r, int(k[1:])] = pc
            out[f"{name}/tr_pairs0"] = pairs
            print("frames", name, per_k[:, 0], tper_k[:, 0] if F > 1 else None, flush=True)
    np.savez_compressed(OUT / "frames.npz", **out)


def gen_transition_quirks():
    """Adversarial index-level cases for the order-dependent bookkeeping
    (EU:278-318).  The reference function is driven with hand-made vectors:
    tile centres themselves, so nearest tile == the chosen index."""
    vet, U = _ref()
    rng = np.random.default_rng(21)
    out = {}
    for ci, (n, users, spread) in enumerate([(20, 12, 3), (20, 60, 21), (50, 200, 6), (200, 64, 201), (20, 2, 1), (20, 1, 1), (50, 500, 51)]):
        L = U.generate_fibonacci_lattice(n)
        T = len(L)
        for rep in range(6):
            p = rng.integers(0, min(spread, T), size=users)
            c = rng.integers(0, min(spread, T), size=users)
            if rep % 2 == 1:
                c = np.where(rng.uniform(size=users) < 0.7, p, c)
            prior = {f"u{u}": L[p[u]] for u in range(users)}
            cur = {f"u{u}": L[c[u]] for u in range(users)}
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                e, wts, asg = U.compute_transition_entropy(prior, cur, L, U.EntropyConfig(), 120)
            key = f"q{ci}_{rep}"
            out[key + "/T"] = np.array(T)
            out[key + "/p"] = p.astype(np.int32)
            out[key + "/c"] = c.astype(np.int32)
            out[key + "/e"] = np.array(float(e))
            assert all(asg[f"u{u}"] == (p[u], c[u]) for u in range(users))
    # SURVEY Appendix B quirk exerciser
    L = U.generate_fibonacci_lattice(200)
    p = np.full(12, 100)
    c = np.array([100, 100, 92, 92] * 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e, _, _ = U.compute_transition_entropy({f"u{u}": L[p[u]] for u in range(12)}, {f"u{u}": L[c[u]] for u in range(12)}, L, U.EntropyConfig(), 120)
    out["appB/T"] = np.array(201); out["appB/p"] = p.astype(np.int32); out["appB/c"] = c.astype(np.int32); out["appB/e"] = np.array(float(e))
    np.savez_compressed(OUT / "transition_quirks.npz", **out)
    print("quirks", len(out) // 4, "appB", float(e))


def gen_analyzers():
    """End-to-end run_analysis-style fixtures: CSV directory -> process_directory
    -> compute_entropy for both analyzers (SA:68-164, TA:68-175), including the
    ragged directory of SURVEY Appendix B."""
    import pandas as pd
    vet, U = _ref()
    out = {}

    def run(dirpath, tile_counts, ecfg, tag, order):
        # the reference globs in OS order (SA:85); force a known order by patching Path.glob
        real_glob = Path.glob

        def sorted_glob(self, pat):
            return iter([Path(dirpath) / f"{n}.csv" for n in order])
        Path.glob = sorted_glob
        try:
            with tempfile.TemporaryDirectory() as od:
                cfg = vet.AnalyzerConfig(video_width=W0, video_height=H0, tile_counts=tile_counts,
                                         output_dir=Path(od), entropy_config=ecfg)
                sa = vet.SpatialEntropyAnalyzer(cfg)
                sa.process_directory(Path(dirpath))
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    df = sa.compute_entropy()
                ta = vet.TransitionEntropyAnalyzer(cfg)
                ta.process_directory(Path(dirpath))
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    tdf = ta.compute_entropy()
        finally:
            Path.glob = real_glob
        L0 = U.generate_fibonacci_lattice(tile_counts[0])
        idx = {c: i for i, c in enumerate(L0)}
        out[f"{tag}/order"] = np.array(order)
        out[f"{tag}/tile_counts"] = np.array(tile_counts)
        out[f"{tag}/sp_time"] = df["time"].to_numpy(dtype=np.float64)
        out[f"{tag}/sp_entropy"] = df["entropy"].to_numpy(dtype=np.float64)
        hist = np.zeros((len(df), len(L0)))
        asg = np.full((len(df), len(order)), 0xFFFF, dtype=np.uint16)
        for r in range(len(df)):
            for c, w in df["tile_weights"][r].items():
                hist[r, idx[c]] = w
            for k, t in df["tile_assignments"][r].items():
                asg[r, order.index(k)] = t
        out[f"{tag}/sp_hist0"] = hist
        out[f"{tag}/sp_assign0"] = asg
        out[f"{tag}/tr_time"] = tdf["time"].to_numpy(dtype=np.float64)
        out[f"{tag}/tr_entropy"] = tdf["entropy"].to_numpy(dtype=np.float64)
        pm = np.zeros((len(tdf), len(L0)), dtype=np.int32)
        for r in range(len(tdf)):
            for c, w in tdf["tile_weights"][r].items():
                pm[r, idx[c]] = w
        out[f"{tag}/tr_prev_count0"] = pm
        print("analyzer", tag, df["entropy"].to_numpy()[:4], tdf["entropy"].to_numpy()[:4], flush=True)

    # (1) ragged directory, SURVEY Appendix B
    with tempfile.TemporaryDirectory() as d:
        pd.DataFrame({"time": [5.0, 5.1, 5.2, 5.3], "2dmu": [.5, .5, .6, .7], "2dmv": [.5] * 4}).to_csv(f"{d}/a.csv", index=False)
        pd.DataFrame({"time": [9.0, 9.1, 9.14, 9.3], "2dmu": [.1, .2, .3, .4], "2dmv": [.2] * 4}).to_csv(f"{d}/b.csv", index=False)
        pd.DataFrame({"time": [1.2, 1.0, 1.1], "2dmu": [.9] * 3, "2dmv": [.9] * 3}).to_csv(f"{d}/c.csv", index=False)
        run(d, [20], U.EntropyConfig(), "ragged", ["a", "b", "c"])
    # (2) synthetic directory: 10 users x 30 frames at 10 Hz, fp64 CSV values with extra columns + a NaN row
    rng = np.random.default_rng(33)
    with tempfile.TemporaryDirectory() as d:
        names = [f"user{u:02d}" for u in range(10)]
        csvs = {}
        for u, nme in enumerate(names):
            F = 30 if u % 3 else 27
            t = 100.0 + u + np.arange(F) * 0.1 + rng.uniform(-0.02, 0.02, size=F)
            mu = np.clip(0.5 + np.cumsum(rng.normal(0, 0.02, size=F)), 0, 1)
            mv = np.clip(0.5 + np.cumsum(rng.normal(0, 0.01, size=F)), 0, 1)
            df = pd.DataFrame({"frame": np.arange(F), "time": t, "2dmu": mu, "2dmv": mv, "other": 1.0})
            if u == 4:
                df.loc[5, "2dmu"] = np.nan
            df.to_csv(f"{d}/{nme}.csv", index=False)
            csvs[nme] = df
        run(d, [20, 50], U.EntropyConfig(), "dir10_default", names)
        run(d, [50, 20, 200], U.EntropyConfig(fov_angle=90.0, use_weight_distribution=False), "dir10_unw", names[::-1])
        # keep the CSV payload so the GPU box can rebuild the same directory
        for nme, df in csvs.items():
            out[f"dir10/{nme}"] = df[["time", "2dmu", "2dmv"]].to_numpy(dtype=np.float64)
    np.savez_compressed(OUT / "analyzers.npz", **out)

NAIVE_TILES = [(30, 30), (45, 90), (10, 20), (360, 180), (3, 3), (120, 60)]   # (tile_width, tile_height) degrees


def gen_naive():
    """Latitude/longitude grid tiling (NaiveSpatialEntropyAnalyzer NA:39-241, EU:335-453): the reference's
    tile key of EVERY reachable cell for several tile sizes, compute_naive_spatial_entropy on seeded frames
    (both normalisations, missing users), and the analyzer end to end on a CSV directory."""
    import pandas as pd
    vet, U = _ref()
    out = {}
    # (1) every cell of the default grid through the reference decode chain and find_naive_tile_index
    lon = np.zeros((H0 + 1, W0 + 1))
    lat = np.zeros((H0 + 1, W0 + 1))
    for py in range(H0 + 1):
        for px in range(W0 + 1):
            lon[py, px], lat[py, px], _ = ref_cell_vector(px, py, W0, H0)
    for tw, th in NAIVE_TILES:
        li = np.zeros((H0 + 1, W0 + 1), dtype=np.int32)
        la = np.zeros((H0 + 1, W0 + 1), dtype=np.int32)
        for py in range(H0 + 1):
            for px in range(W0 + 1):
                key = U.find_naive_tile_index(vet.RadialPoint(lon=lon[py, px], lat=lat[py, px]), th, tw)
                a, b = key.split("_")
                li[py, px], la[py, px] = int(a), int(b)
        out[f"cells/{tw}x{th}/lon_idx"] = li
        out[f"cells/{tw}x{th}/lat_idx"] = la
    # (2) seeded frames
    cases = [("n_30_w", 5, 400, (30, 30), True, 0.0, False), ("n_30_u", 5, 400, (30, 30), False, 0.1, False),
             ("n_45x90_u", 4, 3, (45, 90), False, 0.0, True), ("n_10x20_w", 3, 1500, (10, 20), True, 0.05, True),
             ("n_360_u", 3, 50, (360, 180), False, 0.0, True), ("n_3_u", 3, 6000, (3, 3), False, 0.0, True),
             ("n_120x60_one", 2, 1, (120, 60), False, 0.0, False)]
    for tag, F, Un, (tw, th), use_w, missing, iid in cases:
        packed = synth_packed(F, Un, 9000 + Un, missing=missing, iid=iid)
        ent = np.zeros(F)
        keys = np.full((F, Un, 2), -1, dtype=np.int32)
        nkeys = np.zeros(F, dtype=np.int64)
        for f in range(F):
            mu = packed[f, :, 1].astype(np.float64)
            mv = packed[f, :, 2].astype(np.float64)
            ok = ~(np.isnan(mu) | np.isnan(mv))
            px = U.normalize_to_pixel(np.where(ok, mu, 0.0), W0)
            py = U.normalize_to_pixel(np.where(ok, mv, 0.0), H0)
            pts = {}
            for u in range(Un):
                if ok[u]:
                    lo, la_, _ = ref_cell_vector(px[u], py[u], W0, H0)
                    pts[f"u{u:05d}"] = vet.RadialPoint(lon=lo, lat=la_)
                else:
                    pts[f"u{u:05d}"] = None
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                e, wts, asg = U.compute_naive_spatial_entropy(pts, th, tw, U.EntropyConfig(use_weight_distribution=use_w))
            ent[f] = e
            nkeys[f] = len(wts)
            assert sum(wts.values()) == ok.sum()
            for u in range(Un):
                if ok[u]:
                    a, b = asg[f"u{u:05d}"].split("_")
                    keys[f, u] = (int(a), int(b))
        out[f"frames/{tag}/packed"] = packed
        out[f"frames/{tag}/tile"] = np.array([tw, th])
        out[f"frames/{tag}/use_w"] = np.array(use_w)
        out[f"frames/{tag}/entropy"] = ent
        out[f"frames/{tag}/keys"] = keys
        out[f"frames/{tag}/nkeys"] = nkeys
        print("naive", tag, ent, flush=True)
    # (3) analyzer end to end on the dir10 directory of analyzers.npz
    src = np.load(OUT / "analyzers.npz")
    names = [f"user{u:02d}" for u in range(10)]
    real_glob = Path.glob
    with tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as od:
        for nme in names:
            a = src[f"dir10/{nme}"]
            pd.DataFrame({"time": a[:, 0], "2dmu": a[:, 1], "2dmv": a[:, 2]}).to_csv(f"{d}/{nme}.csv", index=False)
        Path.glob = lambda self, pat: iter([Path(d) / f"{n}.csv" for n in names])
        try:
            for tw, th, use_w in ((30, 30, True), (45, 90, False)):
                from viewport_entropy_toolkit.config import NaiveAnalyzerConfig
                cfg = NaiveAnalyzerConfig(video_width=W0, video_height=H0, output_dir=Path(od), tile_width=tw, tile_height=th,
                                              entropy_config=U.EntropyConfig(use_weight_distribution=use_w))
                na = vet.NaiveSpatialEntropyAnalyzer(cfg)
                na.process_directory(Path(d))
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    df = na.compute_entropy()
                out[f"analyzer/{tw}x{th}_{int(use_w)}/time"] = df["time"].to_numpy(dtype=np.float64)
                out[f"analyzer/{tw}x{th}_{int(use_w)}/entropy"] = df["entropy"].to_numpy(dtype=np.float64)
                assert df["tile_weights"].isna().all() and df["tile_assignments"].isna().all()
                print("naive analyzer", tw, th, use_w, df["entropy"].to_numpy()[:4], flush=True)
        finally:
            Path.glob = real_glob
    np.savez_compressed(OUT / "naive.npz", **out)


def gen_geometry():
    """Tile geometry for renders (DU:58-225, 412-743): boundary segments of the Fibonacci tiles, latitude /
    longitude tile boxes and the tile areas, from the reference's own functions."""
    load_reference()
    from viewport_entropy_toolkit.utilities import data_utils as DU
    out = {}
    for n in (3, 20, 50, 200):
        b = DU.get_fb_tile_boundaries(n)
        rows = [(i, e, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z) for i, edges in b.items() for e, (p1, p2) in enumerate(edges)]
        out[f"fb{n}/edges"] = np.array(rows, dtype=np.float64).reshape(-1, 8)
        out[f"fb{n}/tiles"] = np.array([len(b)])
        if n in (20, 50):
            area, frac = DU.compute_fb_tile_areas(n)
            out[f"fb{n}/area"] = np.array([area[i] for i in range(len(b))])
            out[f"fb{n}/fraction"] = np.array([frac[i] for i in range(len(b))])
        print("geometry fb", n, len(rows), flush=True)
    for nh, nv in ((4, 2), (12, 6)):
        t = DU.get_lat_lon_tiles(nh, nv)
        keys = sorted(t)
        rows = [(ki, e, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z) for ki, k in enumerate(keys) for e, (p1, p2) in enumerate(t[k])]
        out[f"ll{nh}x{nv}/keys"] = np.array(keys)
        out[f"ll{nh}x{nv}/edges"] = np.array(rows, dtype=np.float64)
        area, frac = DU.compute_lat_lon_tile_areas(nh, nv)
        out[f"ll{nh}x{nv}/area"] = np.array([area[k] for k in keys])
    np.savez_compressed(OUT / "geometry.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["lattices", "decode", "nearest", "weights", "frames", "quirks", "analyzers", "naive", "geometry"]
    for w in which:
        {"lattices": gen_lattices, "decode": gen_decode, "nearest": gen_nearest, "weights": gen_weights,
         "frames": gen_frames, "quirks": gen_transition_quirks, "analyzers": gen_analyzers, "naive": gen_naive, "geometry": gen_geometry}[w]()
