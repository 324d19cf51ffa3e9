"""Value types of the viewport-entropy API.

Same names, fields and validation as the reference's data_types.py
(Point DT:30-58, RadialPoint DT:61-103, Vector DT:106-216, errors DT:20-27) so
that user code written against the reference keeps working.  These objects only
exist at the Python boundary; the device path works on packed tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Union, Tuple

import numpy as np


class SpatialError(Exception):
    """Root of the package's exception hierarchy (DT:20-22)."""


class ValidationError(SpatialError):
    """Invalid input data or configuration (DT:25-27)."""


@dataclass(frozen=True)
class Point:
    """Pixel coordinates of a viewport centre; negative values are rejected (DT:47-50)."""
    pixel_x: int
    pixel_y: int

    def __post_init__(self) -> None:
        if self.pixel_x < 0 or self.pixel_y < 0:
            raise ValidationError("Pixel coordinates cannot be negative")

    def as_tuple(self) -> Tuple[int, int]:
        return (self.pixel_x, self.pixel_y)


@dataclass(frozen=True)
class RadialPoint:
    """Longitude/latitude in degrees, lon in [-180,180], lat in [-90,90] (DT:78-83)."""
    lon: float
    lat: float

    def __post_init__(self) -> None:
        if not -180 <= self.lon <= 180:
            raise ValidationError("Longitude must be between -180 and 180 degrees")
        if not -90 <= self.lat <= 90:
            raise ValidationError("Latitude must be between -90 and 90 degrees")

    def normalize_coordinates(self) -> "RadialPoint":
        """Wraps into the canonical ranges (DT:93-95)."""
        return RadialPoint(((self.lon + 180) % 360) - 180, ((self.lat + 90) % 180) - 90)

    def as_tuple(self) -> Tuple[float, float]:
        return (self.lon, self.lat)


@dataclass(frozen=True)
class Vector:
    """Cartesian direction; the zero vector is rejected (DT:125-128)."""
    x: float
    y: float
    z: float

    def __post_init__(self) -> None:
        if self.length() == 0:
            raise ValidationError("Vector cannot have zero length")

    def length(self) -> float:
        return np.sqrt(self.x ** 2 + self.y ** 2 + self.z ** 2)

    def normalize(self) -> "Vector":
        n = self.length()
        if n == 0:
            raise ValidationError("Cannot normalize zero-length vector")
        return Vector(x=self.x / n, y=self.y / n, z=self.z / n)

    def dot_product(self, other: "Vector") -> float:
        return self.x * other.x + self.y * other.y + self.z * other.z

    def as_tuple(self) -> Tuple[float, float, float]:
        return (self.x, self.y, self.z)

    def round(self, decimals: int) -> "Vector":
        return Vector(x=np.round(self.x, decimals=decimals), y=np.round(self.y, decimals=decimals),
                      z=np.round(self.z, decimals=decimals))

    @classmethod
    def from_spherical(cls, lon: float, lat: float) -> "Vector":
        """Unit direction of (lon, lat) degrees, each component rounded to six
        decimals exactly as the reference does (DT:198-216)."""
        if not -180 <= lon <= 180:
            raise ValidationError("Longitude must be between -180 and 180 degrees")
        if not -90 <= lat <= 90:
            raise ValidationError("Latitude must be between -90 and 90 degrees")
        from ._tables import spherical_to_vector
        x, y, z = spherical_to_vector(lon, lat)
        return cls(x=float(x), y=float(y), z=float(z))


def convert_vectors_to_coordinates(vectors) -> Tuple[np.ndarray, np.ndarray]:
    """[Vector] -> (longitudes, latitudes) in degrees for plotting (DT:219-276): lon = atan2(y, x),
    lat = asin(z / |v|), longitudes folded into (-180, 180]."""
    if vectors is None or len(vectors) == 0:
        raise ValidationError("Empty vector collection provided")
    try:
        items = vectors.tolist() if isinstance(vectors, np.ndarray) else vectors
        if not all(isinstance(v, Vector) for v in items):
            raise ValidationError("All elements must be Vector instances")
        pts = [RadialPoint(lon=np.degrees(np.arctan2(v.y, v.x)),
                           lat=np.degrees(np.arcsin(v.z / np.sqrt(v.x ** 2 + v.y ** 2 + v.z ** 2)))) for v in items]
        lons = np.array([p.lon for p in pts])
        lats = np.array([p.lat for p in pts])
        lons = np.where(lons > 180, lons - 360, lons)
        lons = np.where(lons <= -180, lons + 360, lons)
        return lons, lats
    except Exception as e:
        raise ValidationError(f"Failed to convert vectors to coordinates: {str(e)}")
