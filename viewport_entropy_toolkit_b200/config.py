"""Configuration dataclasses, field-for-field compatible with the reference
(AnalyzerConfig CFG:39-82, NaiveAnalyzerConfig CFG:84-126, EntropyConfig EU:20-38,
VisualizationConfig VU:32-60, defaults CFG:23-36)."""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Tuple

from .data_types import ValidationError

DEFAULT_VIDEO_DIMENSIONS = {"width": 100, "height": 200}
DEFAULT_TILE_COUNTS = [20, 50, 100, 250, 1000]
DEFAULT_OUTPUT_FORMATS = {"video": ".mp4", "data": ".csv", "plot": ".png"}


@dataclass
class EntropyConfig:
    """fov_angle in degrees (0, 360]; power_factor > 0; use_weight_distribution
    switches between FOV-weighted and nearest-tile histograms (EU:29-38)."""
    fov_angle: float = 120.0
    use_weight_distribution: bool = True
    power_factor: float = 2.0

    def __post_init__(self) -> None:
        if not 0 < self.fov_angle <= 360:
            raise ValidationError("FOV angle must be between 0 and 360 degrees")
        if self.power_factor <= 0:
            raise ValidationError("Power factor must be positive")


@dataclass
class VisualizationConfig:
    """Plot parameters; kept for source compatibility (the rendering stack is
    outside the accelerated path)."""
    figure_size: Tuple[int, int] = (12, 6)
    fov_point_size: int = 10
    tile_point_size: int = 40
    fps: int = 10
    dpi: int = 100

    def __post_init__(self) -> None:
        if any(v <= 0 for v in self.figure_size):
            raise ValidationError("Figure dimensions must be positive")
        for name, msg in (("fov_point_size", "FOV point size"), ("tile_point_size", "Tile point size"),
                          ("fps", "FPS"), ("dpi", "DPI")):
            if getattr(self, name) <= 0:
                raise ValidationError(f"{msg} must be positive")


@dataclass
class AnalyzerConfig:
    """Same fields, defaults and checks as the reference, including the creation
    of output_dir on construction (CFG:62-70) and the shared default tile_counts
    list (CFG:54)."""
    video_width: int = DEFAULT_VIDEO_DIMENSIONS["width"]
    video_height: int = DEFAULT_VIDEO_DIMENSIONS["height"]
    tile_counts: List[int] = field(default_factory=lambda: DEFAULT_TILE_COUNTS)
    output_dir: Path = Path("output")
    entropy_config: EntropyConfig = field(default_factory=EntropyConfig)
    visualization_config: VisualizationConfig = field(default_factory=VisualizationConfig)

    def __post_init__(self) -> None:
        if self.video_width <= 0 or self.video_height <= 0:
            raise ValueError("Video dimensions must be positive")
        if not self.tile_counts:
            raise ValueError("Must specify at least one tile count")
        if any(c <= 0 for c in self.tile_counts):
            raise ValueError("Tile counts must be positive")
        self.output_dir.mkdir(parents=True, exist_ok=True)

    def get_output_path(self, base_name: str, extension: str) -> Path:
        return self.output_dir / f"{base_name}{extension}"


@dataclass
class NaiveAnalyzerConfig:
    """Latitude-longitude grid tiling (CFG:84-126): tile_width / tile_height in degrees.  Same
    fields, defaults (the -1 placeholders) and checks as the reference; whether the sizes divide
    360 / 180 is only checked when an entropy is computed (EU:414-417)."""
    video_width: int = DEFAULT_VIDEO_DIMENSIONS["width"]
    video_height: int = DEFAULT_VIDEO_DIMENSIONS["height"]
    output_dir: Path = Path("output")
    entropy_config: EntropyConfig = field(default_factory=EntropyConfig)
    visualization_config: VisualizationConfig = field(default_factory=VisualizationConfig)
    tile_width: int = -1
    tile_height: int = -1

    def __post_init__(self) -> None:
        if self.video_width <= 0 or self.video_height <= 0:
            raise ValueError("Video dimensions must be positive")
        if not self.tile_height or not self.tile_width:
            raise ValueError("Must specify both tile_height and tile_width")
        self.output_dir.mkdir(parents=True, exist_ok=True)

    def get_output_path(self, base_name: str, extension: str) -> Path:
        return self.output_dir / f"{base_name}{extension}"
