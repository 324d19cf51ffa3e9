"""Tile geometry for renders: boundaries of the Fibonacci-lattice tiles, latitude/longitude tile
boxes, spherical polygon areas (reference: utilities/data_utils.py, DU:58-225 and DU:412-741).

Host-side numpy, like the reference: these functions feed the pyvista tiling renders and the
tile-area tables, never the entropy path, and their cost is O(T x neighbours^2) small-vector
operations -- there is nothing for a GPU to do.  Same names, arguments and results as the
reference (checked value for value against it, tests/golden/geometry.npz); the per-tile work is
organised around one distance matrix and a per-tile cache of great-circle intersections instead
of rebuilding `Vector` objects in every loop.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import _tables
from .data_types import ValidationError, Vector

_NEIGHBOUR_REACH = 1.7   # a tile's neighbours: centres closer than 1.7 x the nearest one (DU:114,133)
_ROUND = 4               # decimals of the length comparisons and of the corner keys (DU:144,171,175; DU:540)


# ---- small vector helpers (DU:412-528) ------------------------------------------------------------
def normalize(v: np.ndarray) -> np.ndarray:
    """v / |v| (DU:412-421)."""
    return v / np.linalg.norm(v)


def find_perpendicular_on_tangent_plane(vec: np.ndarray, midpoint: np.ndarray) -> np.ndarray:
    """Unit vector perpendicular to `vec` in the plane tangent to the sphere at `midpoint` (DU:423-443)."""
    return normalize(np.cross(normalize(midpoint), vec))


def great_circle_intersection(n1: np.ndarray, n2: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """The two antipodal points where the great circles with normals n1, n2 meet (DU:445-467)."""
    p = normalize(np.cross(normalize(n1), normalize(n2)))
    return p, -p


def get_line_segment(v1: Vector, v2: Vector) -> np.ndarray:
    """v1 - v2 as an array (DU:469-481)."""
    return np.array([v1.x - v2.x, v1.y - v2.y, v1.z - v2.z])


def _as_array(v: Vector) -> np.ndarray:
    return np.array([v.x, v.y, v.z])


def find_nearest_point(v1: Vector, v2: Vector, compare_vector: Vector) -> Vector:
    """Whichever of v1, v2 is closer to compare_vector, distances rounded to 4 decimals; v2 on a tie
    (DU:483-501)."""
    d1 = np.linalg.norm(get_line_segment(compare_vector, v1)).round(_ROUND)
    d2 = np.linalg.norm(get_line_segment(compare_vector, v2)).round(_ROUND)
    return v1 if d1 < d2 else v2


def spherical_interpolation(v1: Vector, v2: Vector, t: float) -> np.ndarray:
    """Slerp along the shorter arc (DU:503-528)."""
    p1, p2 = normalize(_as_array(v1)), normalize(_as_array(v2))
    theta = np.arccos(np.clip(np.dot(p1, p2), -1.0, 1.0))
    return (np.sin((1 - t) * theta) * p1 + np.sin(t * theta) * p2) / np.sin(theta)


# ---- tile boundaries (DU:58-225) --------------------------------------------------------------------
def get_fb_tile_boundaries(tile_count: int) -> Dict[int, List[List[Vector]]]:
    """{tile index: [[corner, corner], ...]} for the Fibonacci lattice of `tile_count` (DU:58-189).

    For tile i the neighbours j are the centres within 1.7 x the nearest distance, nearest first.  The
    bisector of (i, j) is the great circle whose normal is c_i - c_j; its crossings with the bisectors of
    the other neighbours k give candidate corners (of each antipodal pair the one nearer to c_i).  The two
    candidates nearest to c_i bound the edge towards j, unless the bisectors of THEIR two neighbours cross
    closer to c_i than the midpoint of (i, j) does -- then j is screened off and contributes no edge."""
    if tile_count <= 0:
        raise ValidationError("Tile counts cannot be less than 1 for to visualize tiling!")
    centres = _tables.fibonacci_lattice(tile_count)            # [T,3], identical to generate_fibonacci_lattice
    T = len(centres)
    boundaries: Dict[int, List[List[Vector]]] = {}
    for i in range(T):
        ci = centres[i]
        normals = ci[None, :] - centres                          # bisector normals c_i - c_j (DU:93-94,108)
        lengths = np.array([np.linalg.norm(normals[j]) for j in range(T)])
        others = np.array([j for j in range(T) if j != i], dtype=np.int64)
        order = others[np.argsort(lengths[others], kind="stable")]  # nearest first, ties in index order
        reach = lengths[order[0]] * _NEIGHBOUR_REACH if len(order) else 0.0
        near = [int(j) for j in order if lengths[j] < reach]

        corner_cache: Dict[Tuple[int, int], Tuple[Vector, float]] = {}

        def corner(j: int, k: int) -> Tuple[Vector, float]:
            """Crossing of bisectors j and k on c_i's side and its rounded distance from c_i."""
            hit = corner_cache.get((j, k))
            if hit is None:
                p1, p2 = great_circle_intersection(normals[j], normals[k])
                d1 = np.linalg.norm(ci - p1).round(_ROUND)
                d2 = np.linalg.norm(ci - p2).round(_ROUND)
                p = p1 if d1 < d2 else p2
                hit = (Vector(p[0], p[1], p[2]), np.linalg.norm(ci - p).round(_ROUND))
                corner_cache[(j, k)] = hit
            return hit

        edges: List[List[Vector]] = []
        for j in near:
            crossings = [(corner(j, k), k) for k in near if k != j]
            if len(crossings) < 2:
                continue
            crossings.sort(key=lambda e: e[0][1])                # stable: ties keep neighbour order
            (first, ka), (second, kb) = crossings[0], crossings[1]
            screen = corner(ka, kb)[1]                           # where the two bounding bisectors cross
            mid = normalize((ci + centres[j]) / 2)
            if screen > np.linalg.norm(ci - mid).round(_ROUND):
                edges.append([first[0], second[0]])
        boundaries[i] = edges
    return boundaries


def get_lat_lon_tiles(num_tiles_horizontal: int, num_tiles_vertical: int, radius: float = 1.0) -> Dict[str, List[List[Vector]]]:
    """{"row_col": edges} of a latitude/longitude grid; the rows touching a pole are triangles (DU:191-225)."""
    lat_step = 180 / num_tiles_vertical
    lon_step = 360 / num_tiles_horizontal
    north, south = Vector(0, 0, radius), Vector(0, 0, -radius)
    tiles: Dict[str, List[List[Vector]]] = {}
    for i in range(num_tiles_vertical):
        lat1, lat2 = -90 + i * lat_step, -90 + (i + 1) * lat_step
        for j in range(num_tiles_horizontal):
            lon1, lon2 = -180 + j * lon_step, -180 + (j + 1) * lon_step
            bl = Vector.from_spherical(lat=lat1, lon=lon1)
            br = Vector.from_spherical(lat=lat1, lon=lon2)
            tr = Vector.from_spherical(lat=lat2, lon=lon2)
            tl = Vector.from_spherical(lat=lat2, lon=lon1)
            if lat2 >= 90:
                edges = [[bl, br], [bl, north], [br, north]]
            elif lat1 <= -90:
                edges = [[tr, tl], [tr, south], [tl, south]]
            else:
                edges = [[bl, br], [bl, tl], [br, tr], [tr, tl]]
            tiles[f"{i}_{j}"] = edges
    return tiles


# ---- corners and areas (DU:530-741) --------------------------------------------------------------------
def get_tile_corners(tile_boundaries: List[List[Vector]]) -> List[Vector]:
    """Corners in edge order: consecutive corners share an edge (DU:530-575).  Corners are identified by
    their coordinates rounded to 4 decimals."""
    first, second = (p.round(decimals=_ROUND) for p in tile_boundaries[0])
    walk = [first, second]
    visited = {first: True, second: True}
    adjacent: Dict[Vector, List[Vector]] = {}
    for a, b in tile_boundaries:
        a, b = a.round(decimals=_ROUND), b.round(decimals=_ROUND)
        adjacent.setdefault(a, []).append(b)
        adjacent.setdefault(b, []).append(a)
        visited.setdefault(a, False)
        visited.setdefault(b, False)
    here = second
    while not visited[adjacent[here][0]] or not visited[adjacent[here][1]]:
        here = adjacent[here][0] if not visited[adjacent[here][0]] else adjacent[here][1]
        walk.append(here)
        visited[here] = True
    return walk


def triangulate_spherical_polygon(tile_corners: List[Vector]) -> List[List[Vector]]:
    """Fan triangulation from the first corner (DU:577-600)."""
    if len(tile_corners) < 3:
        raise ValueError("At least 3 boundary points are needed for a polygon.")
    anchor = tile_corners[0]
    return [[anchor, tile_corners[i], tile_corners[i + 1]] for i in range(1, len(tile_corners) - 1)]


def angle_at_vertex(v1: np.ndarray, v2: np.ndarray, v3: np.ndarray) -> float:
    """Angle at v1 between the great circles (v1, v2) and (v1, v3) (DU:602-621)."""
    t1 = v2 - np.dot(v2, v1) * v1
    t2 = v3 - np.dot(v3, v1) * v1
    t1 /= np.linalg.norm(t1)
    t2 /= np.linalg.norm(t2)
    return np.arccos(np.clip(np.dot(t1, t2), -1.0, 1.0))


def calculate_spherical_triangle_area(P1: Vector, P2: Vector, P3: Vector, radius: float = 1.0) -> float:
    """Spherical excess x radius^2 (DU:623-656)."""
    a, b, c = (normalize(_as_array(p)) for p in (P1, P2, P3))
    excess = angle_at_vertex(a, b, c) + angle_at_vertex(b, c, a) + angle_at_vertex(c, a, b) - np.pi
    return excess * (radius ** 2)


def compute_spherical_polygon_area(tile_boundaries: List[List[Vector]], radius=1.0) -> float:
    """Sum of the fan triangles' areas (DU:658-680)."""
    total = 0.0
    for p1, p2, p3 in triangulate_spherical_polygon(get_tile_corners(tile_boundaries)):
        total += calculate_spherical_triangle_area(p1, p2, p3, radius)
    return total


def _areas(tiles: Dict) -> Tuple[Dict, Dict]:
    area = {key: compute_spherical_polygon_area(edges) for key, edges in tiles.items()}
    whole = 4 * np.pi
    return area, {key: a / whole for key, a in area.items()}


def compute_fb_tile_areas(tile_count: int) -> Tuple[Dict[int, float], Dict[int, float]]:
    """(tile_area_dict, fraction_of_sphere_dict) of the Fibonacci tiles (DU:682-710)."""
    if tile_count <= 0:
        raise ValidationError("Number of points must be positive!")
    return _areas(get_fb_tile_boundaries(tile_count))


def compute_lat_lon_tile_areas(num_tiles_horizontal: int, num_tiles_vertical: int) -> Tuple[Dict[str, float], Dict[str, float]]:
    """(tile_area_dict, fraction_of_sphere_dict) of the latitude/longitude tiles (DU:712-743)."""
    if num_tiles_horizontal <= 0 or num_tiles_vertical <= 0:
        raise ValidationError("Number of tiles horizontal and vertical must be positive!")
    return _areas(get_lat_lon_tiles(num_tiles_horizontal, num_tiles_vertical))
