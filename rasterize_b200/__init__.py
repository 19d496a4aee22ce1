"""rasterize_b200 — B200-native fill pipeline (flatten -> signed-difference raster -> paint/composite)
behind the reference's `Rasterizer` interface.  See DESIGN.md and include/rasterize_b200.h."""
from .api import (PARSE_INFO_DTYPE, Align, DEFAULT_FLATNESS, JOB_DTYPE, DevicePath, DevicePathBatch, MultiGpuRasterizer, PathBatch, PreparedBatch, FillRule, GpuRasterizer, GradLinear, GradRadial, GradSpread, GradStop, Job,
                  LinColor, LineCap, LineJoin, Path, PathBuilder, RgpuError, Size, StrokeStyle, Transform, Units, paint_from_desc)
from . import ffi

__all__ = ["PARSE_INFO_DTYPE", "Align", "DEFAULT_FLATNESS", "JOB_DTYPE", "DevicePath", "DevicePathBatch", "MultiGpuRasterizer", "PathBatch", "PreparedBatch", "FillRule", "GpuRasterizer", "GradLinear", "GradRadial", "GradSpread", "GradStop",
           "Job", "LinColor", "LineCap", "LineJoin", "StrokeStyle", "Path", "PathBuilder", "RgpuError", "Size", "Transform", "Units", "paint_from_desc", "ffi"]
