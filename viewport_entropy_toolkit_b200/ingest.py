"""Directory ingest: VCT CSV files -> packed [F, U, 3] float64 tensor.

Vectorised restatement of process_viewport_data (DU:289-342) and
format_trajectory_data (DU:345-410): per file read (time, 2dmu, 2dmv), drop NaN
rows, shift time to start at 0, validate the [0,1] range; then bin time to 0.1 s
(DU:371), order frames by FIRST APPEARANCE while scanning files in the given
order and rows in file order (DU:377-387), and let a later sample of the same
user in the same bin overwrite the earlier one (DU:400-404).  Absent samples are
NaN (the reference's None).  The reference's O(F^2 U) list searches become one
pandas `unique` + indexer lookup.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Sequence, Tuple, Union

import numpy as np
import pandas as pd

from .data_types import RadialPoint, ValidationError, Vector

COLUMNS = ["time", "2dmu", "2dmv"]


def read_viewport_csv(filepath: Union[str, Path]) -> Tuple[pd.DataFrame, str]:
    """One trajectory file -> (frame with time/2dmu/2dmv, identifier = file stem)."""
    filepath = Path(filepath)
    try:
        if not filepath.exists():
            raise FileNotFoundError(f"File not found: {filepath}")
        data = pd.read_csv(filepath, usecols=COLUMNS).dropna()
        if data.empty:
            raise ValidationError(f"No valid data found in {filepath}")
        data = data[COLUMNS].astype(np.float64)
        data["time"] = data["time"] - data["time"].min()
        for col in ("2dmu", "2dmv"):
            v = data[col].to_numpy()
            if np.any((v < 0) | (v > 1)):
                raise ValidationError("Normalized coordinates must be between 0 and 1")
        return data, filepath.stem
    except Exception as e:  # the reference wraps everything, FileNotFoundError included (DU:341-342)
        raise ValidationError(f"Error processing viewport data: {str(e)}")


def pack_trajectories(trajectories: Sequence[Tuple[str, pd.DataFrame]]) -> Tuple[np.ndarray, np.ndarray, List[str]]:
    """[(identifier, frame)] -> (packed[F,U,3] float64, times[F], identifiers)."""
    if not trajectories:
        raise ValidationError("No trajectory data provided")
    identifiers = [name for name, _ in trajectories]
    binned = [df["time"].round(1).to_numpy(dtype=np.float64) for _, df in trajectories]
    times = pd.unique(np.concatenate(binned))  # first-appearance order
    index = pd.Index(times)
    F, U = len(times), len(trajectories)
    packed = np.full((F, U, 3), np.nan, dtype=np.float64)
    packed[:, :, 0] = times[:, None]
    # a duplicated identifier shares one column in the reference (dict keyed by name); last file wins per bin
    column = {}
    for u, name in enumerate(identifiers):
        column.setdefault(name, u)
    for (name, df), t in zip(trajectories, binned):
        rows = index.get_indexer(t)
        keep = ~pd.Series(rows).duplicated(keep="last").to_numpy()  # later sample of a bin overwrites
        u = column[name]
        packed[rows[keep], u, 1] = df["2dmu"].to_numpy(dtype=np.float64)[keep]
        packed[rows[keep], u, 2] = df["2dmv"].to_numpy(dtype=np.float64)[keep]
    return packed, times, identifiers


def list_csv_files(directory: Union[str, Path]) -> List[Path]:
    """Files in the order the reference would visit them: Path.glob order (SA:85)."""
    return list(Path(directory).glob("*.csv"))


def load_directory(directory: Union[str, Path], order: Sequence[str] = None):
    """directory -> (packed, times, identifiers).  `order` (file stems) pins the user
    order explicitly; by default it is the filesystem's glob order like the reference."""
    directory = Path(directory)
    files = list_csv_files(directory) if order is None else [directory / f"{stem}.csv" for stem in order]
    trajectories = []
    for f in files:
        df, name = read_viewport_csv(f)
        trajectories.append((name, df))
    return pack_trajectories(trajectories)


# ---- the reference's own function names (DU:289-410), for code written against them --------------
def process_viewport_data(filepath: Union[str, Path], video_width: int, video_height: int) -> Tuple[pd.DataFrame, str]:
    """process_viewport_data (DU:289-342): one CSV -> (DataFrame[time, 2dmu, 2dmv, pixel_x, pixel_y, lon, lat],
    identifier).  Same arithmetic, vectorised: pixel = trunc(value * dim) (DU:261), lon = (px/W)*360-180,
    lat = 90-(py/H)*180 (DU:283-284).  Every failure is re-raised as ValidationError like DU:341-342."""
    try:
        filepath = Path(filepath)
        if not filepath.exists():
            raise FileNotFoundError(f"File not found: {filepath}")
        data = pd.read_csv(filepath, usecols=COLUMNS).dropna()
        if data.empty:
            raise ValidationError(f"No valid data found in {filepath}")
        data["time"] -= data["time"].min()
        for col, dim, out in (("2dmu", video_width, "pixel_x"), ("2dmv", video_height, "pixel_y")):
            v = data[col].to_numpy()
            if np.any(v < 0) or np.any(v > 1):
                raise ValidationError("Normalized coordinates must be between 0 and 1")  # DU:256-257
            data[out] = (v * dim).astype(int)
        if video_width <= 0 or video_height <= 0:
            raise ValidationError("Video dimensions must be positive")                     # DU:236-237
        if video_width % 2 != 0 or video_height % 2 != 0:
            raise ValidationError("Video dimensions must be even numbers")                 # DU:238-240
        px = data["pixel_x"].to_numpy(dtype=np.float64)
        py = data["pixel_y"].to_numpy(dtype=np.float64)
        data["lon"] = (px / video_width) * 360 - 180
        data["lat"] = 90 - (py / video_height) * 180
        return data, filepath.stem
    except Exception as e:
        raise ValidationError(f"Error processing viewport data: {str(e)}")


def format_trajectory_data(trajectory_data: List[Tuple[str, pd.DataFrame]]) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """format_trajectory_data (DU:345-410): [(identifier, data)] -> (points_df, vectors_df), object frames with a
    `time` column and one column per identifier holding RadialPoint / Vector or None.  Frames are ordered by first
    appearance of their 0.1 s bin (DU:371-387), a later sample of a bin overwrites (DU:400-404), lon/lat are rounded
    with Python's round(., 1) and wrapped like DU:390-397; like the reference, `data["time"]` is rounded in place.
    The reference's per-row list searches (O(F^2 U)) are replaced by one indexer lookup."""
    if not trajectory_data:
        raise ValidationError("No trajectory data provided")
    identifiers = [pair[0] for pair in trajectory_data]
    for _, data in trajectory_data:
        data["time"] = data["time"].round(1)
    times = pd.unique(np.concatenate([d["time"].to_numpy(dtype=np.float64) for _, d in trajectory_data]))
    index = pd.Index(times)
    F = len(times)
    points = {"time": list(times)}
    vectors = {"time": list(times)}
    for name in identifiers:
        points.setdefault(name, [None] * F)
        vectors.setdefault(name, [None] * F)
    for name, data in trajectory_data:
        rows = index.get_indexer(data["time"].to_numpy(dtype=np.float64))
        for r, lon, lat in zip(rows, data["lon"].tolist(), data["lat"].tolist()):
            lon = round(lon, 1)
            lat = round(lat, 1)
            if lon <= -180:
                lon = (lon + 360) % 360 - 180
            if lat <= -90:
                lat = (lat + 180) % 180 - 90
            points[name][r] = RadialPoint(lon=lon, lat=lat)
            vectors[name][r] = Vector.from_spherical(lon, lat)
    return pd.DataFrame(points), pd.DataFrame(vectors)
