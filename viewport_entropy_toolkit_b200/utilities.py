"""Functional API with the reference's names (utilities/entropy_utils.py,
utilities/data_utils.py), backed by the device engine.

These functions take arbitrary `tile_centers` lists and dicts of `Vector`, like
the reference; each call builds (or reuses) an engine for that centre list and
runs the direct per-sample kernels.  They exist for source compatibility and
validation -- at scale use the analyzers' packed-tensor entry points.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _tables
from .config import EntropyConfig, DEFAULT_VIDEO_DIMENSIONS
from .data_types import Point, RadialPoint, ValidationError, Vector
from .engine import get_engine
from .ingest import format_trajectory_data, process_viewport_data  # noqa: F401  (reference names, DU:289-410)
from .geometry import (  # noqa: F401  (render support, DU:58-225, 412-743)
    angle_at_vertex, calculate_spherical_triangle_area, compute_fb_tile_areas, compute_lat_lon_tile_areas,
    compute_spherical_polygon_area, find_nearest_point, find_perpendicular_on_tangent_plane, get_fb_tile_boundaries,
    get_lat_lon_tiles, get_line_segment, get_tile_corners, great_circle_intersection, normalize,
    spherical_interpolation, triangulate_spherical_polygon)

_TINY_VIDEO = (2, 2)  # the vector entry points do not use the video grid


def _centres_array(tile_centers: Sequence[Vector]) -> np.ndarray:
    if not tile_centers:
        raise ValidationError("No tile centers provided")
    return np.array([[c.x, c.y, c.z] for c in tile_centers], dtype=np.float64)


def _engine_for(tile_centers: Sequence[Vector], config: Optional[EntropyConfig] = None):
    c = _centres_array(tile_centers)
    return get_engine(_TINY_VIDEO[0], _TINY_VIDEO[1], [len(c)], config or EntropyConfig(), centres=[c])


def _vec(v: Vector) -> List[float]:
    return [v.x, v.y, v.z]


def generate_fibonacci_lattice(num_points: int) -> List[Vector]:
    """DU:25-56: 2*int(n/2)+1 golden-angle lattice points."""
    if num_points <= 0:
        raise ValidationError("Number of points must be positive")
    return [Vector(float(x), float(y), float(z)) for x, y, z in _tables.fibonacci_lattice(num_points)]


def validate_video_dimensions(width: int, height: int) -> None:
    """DU:227-240."""
    if width <= 0 or height <= 0:
        raise ValidationError("Video dimensions must be positive")
    if width % 2 != 0 or height % 2 != 0:
        raise ValidationError("Video dimensions must be even numbers")


def normalize_to_pixel(normalized: np.ndarray, dimension: int) -> np.ndarray:
    """DU:243-261 (ingest side, host): trunc(normalized * dimension), [0,1] enforced."""
    normalized = np.asarray(normalized)
    if np.any((normalized < 0) | (normalized > 1)):
        raise ValidationError("Normalized coordinates must be between 0 and 1")
    if dimension <= 0:
        raise ValidationError("Dimension must be positive")
    return (normalized * dimension).astype(int)


def pixel_to_spherical(point: Point, video_width: int, video_height: int) -> RadialPoint:
    """DU:264-286."""
    validate_video_dimensions(video_width, video_height)
    if point.pixel_x > video_width or point.pixel_y > video_height:
        raise ValidationError("Pixel coordinates exceed video dimensions")
    return RadialPoint(lon=(point.pixel_x / video_width) * 360 - 180, lat=90 - (point.pixel_y / video_height) * 180)


def vector_angle_distance(v1: Vector, v2: Vector) -> float:
    """EU:41-67: angle between two vectors in radians."""
    # one table-free handle serves every pair (vet_vector_angles): no engine per distinct v2
    eng = get_engine(_TINY_VIDEO[0], _TINY_VIDEO[1], [1], EntropyConfig())
    d = eng.vector_angles(torch.tensor([_vec(v1)], dtype=torch.float64), torch.tensor([_vec(v2)], dtype=torch.float64))
    return float(d[0].item())


def find_angular_distances(vector: Vector, tile_centers: List[Vector]) -> np.ndarray:
    """EU:70-87: array of [tile_index, angular_distance] pairs."""
    eng = _engine_for(tile_centers)
    d = eng.angular_distances(torch.tensor([_vec(vector)], dtype=torch.float64), 0)[0].cpu().numpy()
    return np.stack([np.arange(len(d), dtype=np.float64), d], axis=1)


def find_nearest_tile(vector: Vector, tile_centers: List[Vector]) -> int:
    """EU:89-106: index of the closest tile centre (first one on ties)."""
    eng = _engine_for(tile_centers)
    return int(eng.nearest_tile(torch.tensor([_vec(vector)], dtype=torch.float64), 0)[0].item())


def calculate_tile_weights(vector: Vector, tile_centers: List[Vector], config: EntropyConfig) -> Dict[Vector, float]:
    """EU:108-144: {tile centre: weight} for tiles within fov/2 (or the nearest tile)."""
    eng = _engine_for(tile_centers, config)
    w = eng.tile_weights(torch.tensor([_vec(vector)], dtype=torch.float64), 0)[0].cpu().numpy()
    if config.use_weight_distribution:
        d = eng.angular_distances(torch.tensor([_vec(vector)], dtype=torch.float64), 0)[0].cpu().numpy()
        order = np.argsort(d, kind="stable")  # the reference inserts in order of distance
        return {tile_centers[i]: float(w[i]) for i in order if w[i] > 0}
    return {tile_centers[int(np.argmax(w))]: 1.0}


def compute_spatial_entropy(vector_dict: Dict[str, Optional[Vector]], tile_centers: List[Vector],
                            config: EntropyConfig) -> Tuple[float, Dict[Vector, float], Dict[str, int]]:
    """EU:147-211: (normalised entropy, {tile: weight}, {identifier: nearest tile})."""
    if not vector_dict:
        raise ValidationError("Empty vector dictionary")
    if not tile_centers:
        raise ValidationError("No tile centers provided")
    ids = [k for k, v in vector_dict.items() if v is not None]
    vec = torch.tensor([[_vec(vector_dict[k]) for k in ids]], dtype=torch.float64).reshape(1, len(ids), 3)
    eng = _engine_for(tile_centers, config)
    if not ids:  # every entry None: no weights at all (EU:180-181 skips them)
        vec = torch.full((1, 1, 3), float("nan"), dtype=torch.float64)
    res = eng.spatial_vectors(vec)
    eng.poll_flags()
    hist = res.hist0[0].cpu().numpy()
    assign = res.assign0[0].cpu().numpy()
    weights = {tile_centers[i]: float(hist[i]) for i in np.flatnonzero(hist)}
    return np.float64(res.entropy[0].item()), weights, {k: int(a) for k, a in zip(ids, assign)}


def compute_transition_entropy(prior_vector_dict: dict, current_vector_dict: dict, tile_centers: List[Vector],
                               config: EntropyConfig = None, FOV_angle: float = 120.0,
                               mode: str = "literal") -> Tuple[float, Dict[Vector, int], Dict[str, Tuple[int, int]]]:
    """EU:213-332: (normalised transition entropy, {previous tile: users}, {identifier: (prev, cur)}).
    `config` and `FOV_angle` are accepted and unused, like in the reference (EU:256)."""
    if not prior_vector_dict or not current_vector_dict:
        raise ValidationError("Empty vector dictionary")
    if not tile_centers:
        raise ValidationError("No tile centers provided")
    ids = [k for k in current_vector_dict if k in prior_vector_dict]  # EU:259-261, current-frame order
    if not ids:
        raise ZeroDivisionError("division by zero")  # EU:326
    vec = torch.tensor([[_vec(prior_vector_dict[k]) for k in ids], [_vec(current_vector_dict[k]) for k in ids]],
                       dtype=torch.float64)
    eng = _engine_for(tile_centers)
    res = eng.transition_vectors(vec, mode=mode)
    eng.poll_flags()
    counts = res.prev_count0[0].cpu().numpy()
    pairs = res.pairs0[0].cpu().numpy()
    weights = {tile_centers[i]: int(counts[i]) for i in np.flatnonzero(counts)}
    return np.float64(res.entropy[0].item()), weights, {k: (int(p), int(c)) for k, (p, c) in zip(ids, pairs)}


# ---- latitude/longitude grid tiling (EU:335-453) ------------------------------------------------
def find_naive_tile_index(point: RadialPoint, tile_height: float, tile_width: float) -> str:
    """EU:360-381: "{lon index}_{lat index}" of the grid tile a point sits in (host arithmetic: one
    division per axis, identical to the reference's)."""
    return f"{int((point.lon + 180) / tile_width)}_{int((point.lat + 90) / tile_height)}"


def calculate_naive_tile_weights(point: RadialPoint, tile_height: float, tile_width: float,
                                 config: EntropyConfig) -> Dict[str, float]:
    """EU:335-358: weight 1.0 on the point's own tile."""
    return {find_naive_tile_index(point, tile_height, tile_width): 1.0}


def compute_naive_spatial_entropy(points_dict: Dict[str, Optional[RadialPoint]], tile_height: int, tile_width: int,
                                  config: EntropyConfig) -> Tuple[float, Dict[str, float], Dict[str, str]]:
    """EU:362-453: (normalised entropy, {tile key: users}, {identifier: tile key}) on the device."""
    if not points_dict:
        raise ValidationError("Empty radial points dictionary")
    if not tile_height or not tile_width:
        raise ValidationError("No tile dimensions provided")
    if 180 % tile_height != 0:
        raise ValidationError("Tile height must divide 180!")
    if 360 % tile_width != 0:
        raise ValidationError("Tile width must divide 360!")
    if tile_height < 0 or tile_width < 0:
        raise ValidationError("No tile dimensions provided")  # the reference would index negative tiles
    ids = [k for k, v in points_dict.items() if v is not None]
    num_tiles = int(180.0 / tile_height) * int(360.0 / tile_width)
    if not ids:  # EU:437-448 on an empty histogram
        if config.use_weight_distribution:
            return np.float64(0.0) / (-num_tiles * (1.0 / num_tiles) * np.log2(1.0 / num_tiles)), {}, {}
        raise ZeroDivisionError("float division by zero")
    eng = get_engine(DEFAULT_VIDEO_DIMENSIONS["width"], DEFAULT_VIDEO_DIMENSIONS["height"], [1], config,
                     naive_tiles=(int(tile_width), int(tile_height)))
    lonlat = torch.tensor([[[points_dict[k].lon, points_dict[k].lat] for k in ids]], dtype=torch.float64)
    ent, li, la = eng.naive_points(lonlat, int(tile_width), int(tile_height), config.use_weight_distribution)
    eng.poll_flags()
    keys = [f"{int(a)}_{int(b)}" for a, b in zip(li[0].cpu().numpy(), la[0].cpu().numpy())]
    weights: Dict[str, float] = {}
    for key in keys:
        weights[key] = weights.get(key, 0.0) + 1.0
    return np.float64(ent[0].item()), weights, dict(zip(ids, keys))
