"""Host-side (numpy) construction of the small per-configuration tables that
the device kernels index: Fibonacci lattice centres and the per-axis lon/lat
tables.  Built with the same numpy functions the reference calls, so the values
follow numpy's arcsin/sin/cos bit for bit; uploaded once per handle.

Reference arithmetic: generate_fibonacci_lattice DU:40-54, pixel_to_spherical
DU:283-284, rounding and wrap quirk DU:390-397, Vector.from_spherical DT:204-216.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

GOLDEN = (1 + np.sqrt(5)) / 2


def spherical_to_vector(lon, lat) -> np.ndarray:
    """(lon, lat) degrees -> 6-decimal-rounded direction, broadcast over arrays."""
    theta = np.radians(np.asarray(lon, dtype=np.float64))
    colat = np.radians(90 - np.asarray(lat, dtype=np.float64))
    s = np.sin(colat)
    comps = np.broadcast_arrays(s * np.cos(theta), s * np.sin(theta), np.cos(colat))
    return np.round(np.stack(comps, axis=-1), 6)


def lattice_size(tile_count: int) -> int:
    """A tile_count of n yields 2*int(n/2)+1 lattice points (DU:43-45)."""
    return 2 * int(tile_count / 2) + 1


def fibonacci_lattice(tile_count: int) -> np.ndarray:
    """Centres [T,3] of the golden-angle spiral lattice."""
    half = int(tile_count / 2)
    centres = np.empty((2 * half + 1, 3), dtype=np.float64)
    for row, i in enumerate(range(-half, half + 1)):
        lat = np.arcsin(2 * i / (2 * half + 1)) * 180 / np.pi
        lon = (i % GOLDEN) * 360 / GOLDEN
        lon = ((lon + 180) % 360) - 180
        centres[row] = spherical_to_vector(lon, lat)
    return centres


def axis_tables(width: int, height: int) -> Tuple[np.ndarray, np.ndarray]:
    """Longitude of every pixel column 0..W and latitude of every row 0..H after
    the 0.1-degree decimal rounding (Python round, not rint(x*10)/10) and the
    reference's wrap of lon<=-180 / lat<=-90."""
    lon = np.empty(width + 1, dtype=np.float64)
    lat = np.empty(height + 1, dtype=np.float64)
    for px in range(width + 1):
        v = round(float((np.float64(px) / width) * 360 - 180), 1)
        lon[px] = (v + 360) % 360 - 180 if v <= -180 else v
    for py in range(height + 1):
        v = round(float(90 - (np.float64(py) / height) * 180), 1)
        lat[py] = (v + 180) % 180 - 90 if v <= -90 else v
    return lon, lat
