"""Drives the LIVE reference on packed frames (test infrastructure / CPU-baseline legs of bench.py only).

Everything numeric here is the reference's own code, unmodified, imported through oracle/_refshim.py:
normalize_to_pixel (DU:243-261) -> pixel_to_spherical (DU:264-286) -> rounding / wrap (DU:390-397) ->
Vector.from_spherical (DT:183-216) -> compute_spatial_entropy (EU:147-211) / compute_transition_entropy
(EU:213-332), called per frame like SA:129-161 / TA:129-172 do.  This file only unpacks the packed tensor into
the dicts those functions take.  The product never imports it.
"""
import warnings

import numpy as np

from ._refshim import load_reference, reference_available  # noqa: F401

_lattices = {}


def _ref():
    load_reference()
    import viewport_entropy_toolkit as vet
    from viewport_entropy_toolkit import utilities as U
    return vet, U


def frame_vectors(frame, W, H):
    """frame[U,3] = (time, 2dmu, 2dmv) -> {identifier: Vector} through the reference's decode chain (NaN = absent)."""
    vet, U = _ref()
    mu = frame[:, 1].astype(np.float64)
    mv = frame[:, 2].astype(np.float64)
    ok = ~(np.isnan(mu) | np.isnan(mv))
    px = U.normalize_to_pixel(np.where(ok, mu, 0.0), W)
    py = U.normalize_to_pixel(np.where(ok, mv, 0.0), H)
    d = {}
    for u in np.flatnonzero(ok):
        rp = U.pixel_to_spherical(vet.Point(np.float64(px[u]), np.float64(py[u])), W, H)
        lon, lat = round(float(rp.lon), 1), round(float(rp.lat), 1)   # DU:390-391
        if lon <= -180:
            lon = (lon + 360) % 360 - 180                              # DU:394-395
        if lat <= -90:
            lat = (lat + 180) % 180 - 90                               # DU:396-397
        vet.RadialPoint(lon=lon, lat=lat)                              # DU:399
        d[f"u{u:07d}"] = vet.Vector.from_spherical(lon, lat)           # DU:403
    return d


def _lattice(n):
    if n not in _lattices:
        _lattices[n] = _ref()[1].generate_fibonacci_lattice(n)
    return _lattices[n]


def spatial_frame(frame, W, H, tile_counts, fov, use_w, pf):
    """Mean over tile counts of the reference's compute_spatial_entropy on one frame (SA:142-156)."""
    _, U = _ref()
    d = frame_vectors(frame, W, H)
    cfg = U.EntropyConfig(fov_angle=fov, use_weight_distribution=use_w, power_factor=pf)
    tot = 0.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n in tile_counts:
            tot += U.compute_spatial_entropy(d, _lattice(n), cfg)[0]
    return tot / len(tile_counts)


def transition_pair(prev_frame, cur_frame, W, H, tile_counts):
    """Mean over tile counts of the reference's compute_transition_entropy on one frame pair (TA:148-160)."""
    _, U = _ref()
    prior, cur = frame_vectors(prev_frame, W, H), frame_vectors(cur_frame, W, H)
    tot = 0.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n in tile_counts:
            tot += U.compute_transition_entropy(prior, cur, _lattice(n), U.EntropyConfig(), 120)[0]
    return tot / len(tile_counts)
