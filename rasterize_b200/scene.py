"""Device-resident `Scene::render` (reference src/scene.rs:186-199, 384-459).

The host keeps what the reference keeps on the host — the node table `Pipeline::build` produces (kinds, transforms,
bounding boxes; here loaded from a fixture or built by the caller) and the recursion of `Pipeline::render_rec` — while
every pixel operation runs on the GPU and every layer lives in HBM:

  Fill     -> one FILL job on a window of the current layer (the `view_mut` sub-image of src/scene.rs:412-429); the
              consecutive fills of a layer go to the scene compositor (`rgpu_render_scene`): ONE raster launch in which every
              layer tile stays on its SM while all its fills are blended in order, fused with `Layer::new`'s background
              and, for the root, the RGBA8 export (RGPU_SCENE_ORDERED=1 selects the one-launch-per-fill ordered batch)
  Opacity  -> child rendered into its own device layer, `rgpu_layer_blend_over_dev(.., opacity)`
  Clip     -> clip path rasterized with `Rasterizer::mask` semantics into an f32 device layer, child layer scaled by it
              (`rgpu_layer_scale_by_mask_dev`) and blended over the parent

The only transfer is the final download (LinColor or RGBA8).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import ffi
from .api import FillRule, GpuRasterizer, Job, Path, Transform, paint_from_desc

FILL, GROUP, OPACITY, CLIP = 0, 1, 2, 3


def _as_i32(v: float) -> int:
    """Rust `f64 as i32`: truncation toward zero, saturating."""
    if v != v:
        return 0
    return int(max(-2 ** 31, min(2 ** 31 - 1, math.trunc(v))))


def fill_window(bbox, lx: int, ly: int, W: int, H: int):
    """Window of a W x H layer at (lx, ly) that the Fill arm of `Pipeline::render_rec` draws a node with bounding box `bbox`
    into (src/scene.rs:412-423 + `view_shape`, src/image.rs:588-605): columns [floor(min x) - lx, ceil(max x) - lx + 1),
    rows alike, clamped to the layer.  The reference casts the i32 bounds `as usize`: a NEGATIVE bound (a node that
    starts left of or above the layer — impossible through `Scene::render`, which restricts every bbox to the view, but
    possible for a hand-built node table) wraps to a huge value that `view_shape` clamps to the layer's width / height,
    so such a node gets an EMPTY window and draws nothing.  Reproduced here.  Returns (col_min, row_min, width, height)."""
    def as_usize(v: int) -> int:
        return v if v >= 0 else v + (1 << 64)

    col_min = min(as_usize(_as_i32(math.floor(bbox[0])) - lx), W)
    col_max = min(max(as_usize(_as_i32(math.ceil(bbox[2])) - lx + 1), col_min), W)
    row_min = min(as_usize(_as_i32(math.floor(bbox[1])) - ly), H)
    row_max = min(max(as_usize(_as_i32(math.ceil(bbox[3])) - ly + 1), row_min), H)
    return col_min, row_min, col_max - col_min, row_max - row_min


@dataclass
class PipelineNode:
    """One node of `Pipeline` (src/scene.rs:225-262)."""
    kind: int
    bbox: np.ndarray                     # node.bbox (min x, min y, max x, max y)
    tr: np.ndarray | None = None         # Fill: node transform; Clip: clip transform
    fill_rule: FillRule = FillRule.NonZero
    path: Path | None = None             # Fill path / Clip path
    path_bbox: np.ndarray | None = None  # path.bbox(identity), for objectBoundingBox paints
    paint: object = None
    paint_desc: dict | None = None
    opacity: float = 1.0
    child: int = 0
    children: list = field(default_factory=list)


@dataclass
class Pipeline:
    nodes: list                # children before parents, root last (allocation order of Pipeline::build)
    view: np.ndarray | None    # render view (min x, min y, max x, max y) or None: the root's bbox
    bg: np.ndarray | None

    @staticmethod
    def load_npz(file) -> "Pipeline":
        z = np.load(file)
        nodes = []
        for i in range(int(z["n_nodes"])):
            g = lambda k: z[f"n{i}_{k}"]  # noqa: E731
            n = PipelineNode(kind=int(g("kind")), bbox=g("bbox"), tr=g("tr"), fill_rule=FillRule(int(g("rule"))), opacity=float(g("opacity")),
                             child=int(g("child")), children=[int(c) for c in g("children")])
            if f"n{i}_points" in z:
                n.path = Path(g("points"), g("kinds"), g("sub"), g("closed"))
                n.path_bbox = g("path_bbox")
            if f"n{i}_paint_kind" in z:
                n.paint_desc = dict(kind=int(g("paint_kind")), units=int(g("paint_units")), linear_colors=int(g("paint_linear_colors")),
                                    spread=int(g("paint_spread")), tr=g("paint_tr"), p0=g("paint_p0"), p1=g("paint_p1"),
                                    r0=float(g("paint_r0")), r1=float(g("paint_r1")), solid=g("paint_solid"), stop_pos=g("paint_stop_pos"),
                                    stop_colors=g("paint_stop_colors"))
                n.paint = paint_from_desc(n.paint_desc)
            nodes.append(n)
        return Pipeline(nodes, z["view"] if bool(z["has_view"]) else None, z["bg"] if bool(z["has_bg"]) else None)


class DeviceLayer:
    """`Layer<C>` (src/scene.rs:464-501) in device memory: integer origin, dense row-major pixels."""

    def __init__(self, rast: GpuRasterizer, bbox, channels: int, color=None):
        self.rast = rast
        self.x, self.y = _as_i32(math.floor(bbox[0])), _as_i32(math.floor(bbox[1]))
        self.width = max(0, _as_i32(math.ceil(bbox[2])) - self.x)
        self.height = max(0, _as_i32(math.ceil(bbox[3])) - self.y)
        self.channels = channels
        n = self.width * self.height
        self.ptr = rast.device_alloc(max(1, n * channels * 4))
        # a colour layer is materialised by its first flush (`Layer::new` is fused into the scene kernel)
        self.fresh = channels == 4 and n > 0 and not os.environ.get("RGPU_SCENE_ORDERED")
        self.bg = color
        if not self.fresh:
            if color is not None:
                rast.fill_color(self.ptr, n, color)
            else:
                rast.device_zero(self.ptr, n * channels * 4)
        self.pending: list = []  # FILL jobs not yet submitted (kept alive with their device paths)
        self.keep: list = []

    def free(self):
        if self.ptr:
            self.rast.device_free(self.ptr)
            self.ptr = 0

    def flush(self, rgba_ptr: int = 0):
        """Submit the fills queued on this layer (they may overlap: composited in order).  The synchronous entry point is
        used so that an internal scratch overflow is retried before anything is composited on top.  `rgba_ptr`: device
        RGBA8 image that receives the layer as it stands after these fills."""
        if os.environ.get("RGPU_SCENE_ORDERED") or self.channels != 4:
            if self.pending:
                self.rast.render_batch(self.pending, independent=False, sync=True)
                self.pending = []
            return
        if self.pending or self.fresh or rgba_ptr:
            if self.width and self.height:
                self.rast.render_scene(self.pending, self.ptr, self.width, self.height, fresh=self.fresh, bg=self.bg, rgba_ptr=rgba_ptr, sync=True)
            self.pending = []
            self.fresh = False

    def intersect(self, other: "DeviceLayer"):
        """Intersection rectangle of `Layer::compose` (src/scene.rs:540-549): (self origin, other origin, w, h)."""
        x0, x1 = max(self.x, other.x), min(self.x + self.width, other.x + other.width)
        y0, y1 = max(self.y, other.y), min(self.y + self.height, other.y + other.height)
        if x1 <= x0 or y1 <= y0:
            return None
        return ((y0 - self.y) * self.width + (x0 - self.x), (y0 - other.y) * other.width + (x0 - other.x), x1 - x0, y1 - y0)


def _render_rec(rast: GpuRasterizer, nodes, node_id: int, layer: DeviceLayer, trash: list) -> None:
    node = nodes[node_id]
    if node.kind == FILL:
        # window of the layer covered by the node: `view_shape` clamps like src/image.rs:588-605
        W = layer.width
        col_min, row_min, ww, wh = fill_window(node.bbox, layer.x, layer.y, W, layer.height)
        if ww == 0 or wh == 0:
            return
        align = Transform.new_translate(-math.floor(node.bbox[0]), -math.floor(node.bbox[1]))
        dp = rast.upload(node.path)
        layer.keep.append(dp)
        layer.pending.append(Job(dp, align * Transform.from_array(node.tr), node.fill_rule, ffi.JOB_FILL, layer.ptr, ww, wh, W,
                                 origin=row_min * W + col_min, paint=node.paint, path_bbox=node.path_bbox))
    elif node.kind == GROUP:
        for c in node.children:
            _render_rec(rast, nodes, c, layer, trash)
    elif node.kind == OPACITY:
        child_layer = _render_node(rast, nodes, node.child, None, None, trash)
        layer.flush()  # the fills queued so far come first
        r = layer.intersect(child_layer)
        if r:
            rast.layer_blend_over(layer.ptr, r[0], layer.width, child_layer.ptr, r[1], child_layer.width, r[2], r[3], opacity=node.opacity)
    elif node.kind == CLIP:
        mask_layer = DeviceLayer(rast, node.bbox, 1)
        trash.append(mask_layer)
        align = Transform.new_translate(-float(mask_layer.x), -float(mask_layer.y))
        child_layer = _render_node(rast, nodes, node.child, None, None, trash)
        if mask_layer.width and mask_layer.height:
            dp = rast.upload(node.path)
            mask_layer.keep.append(dp)
            rast.render_batch([Job(dp, align * Transform.from_array(node.tr), node.fill_rule, ffi.JOB_MASK, mask_layer.ptr, mask_layer.width,
                                   mask_layer.height, mask_layer.width)], independent=True, sync=True)
        r = child_layer.intersect(mask_layer)
        if r:
            rast.layer_scale_by_mask(child_layer.ptr, r[0], child_layer.width, mask_layer.ptr, r[1], mask_layer.width, r[2], r[3])
        layer.flush()
        r = layer.intersect(child_layer)
        if r:
            rast.layer_blend_over(layer.ptr, r[0], layer.width, child_layer.ptr, r[1], child_layer.width, r[2], r[3])
    else:
        raise ValueError(f"unknown pipeline node kind {node.kind}")


def _render_node(rast, nodes, node_id, view, bg, trash, flush: bool = True) -> DeviceLayer:
    """`Pipeline::render` (src/scene.rs:384-395): a fresh layer over `view` (or the node's bbox), then the node."""
    layer = DeviceLayer(rast, view if view is not None else nodes[node_id].bbox, 4, bg)
    trash.append(layer)
    _render_rec(rast, nodes, node_id, layer, trash)
    if flush:
        layer.flush()
    return layer


def render(rast: GpuRasterizer, pipeline: Pipeline, rgba: bool = False):
    """Render the pipeline's root.  Returns (x, y, image): LinColor [H, W, 4] f32, or RGBA8 [H, W, 4] u8 with rgba=True.
    Everything stays on the device until the single download at the end."""
    trash: list = []
    try:
        if not pipeline.nodes:
            return 0, 0, np.zeros((0, 0, 4), dtype=np.uint8 if rgba else np.float32)
        root = _render_node(rast, pipeline.nodes, len(pipeline.nodes) - 1, pipeline.view, pipeline.bg, trash, flush=False)
        if rgba and root.width and root.height and not os.environ.get("RGPU_SCENE_ORDERED"):
            # export fused into the root's last fill launch; 4 B / pixel cross PCIe
            n = root.width * root.height
            rgba_ptr = rast.device_alloc(n * 4)
            try:
                root.flush(rgba_ptr)
                img = rast.to_host(rgba_ptr, (root.height, root.width, 4), np.uint8)
            finally:
                rast.device_free(rgba_ptr)
            rast.batch_status()
            return root.x, root.y, img
        root.flush()
        if rgba:
            img = rast.download_rgba8(root.ptr, (root.height, root.width))
        else:
            img = rast.to_host(root.ptr, (root.height, root.width, 4), np.float32)
        rast.batch_status()
        return root.x, root.y, img
    finally:
        rast.sync()
        for l in trash:
            l.free()


def fixture_jobs(rast: GpuRasterizer, sc, layer_ptr: int):
    """FILL jobs of a Fill-only scene fixture (`assets.load_scene`) on a dense layer over the scene's view, exactly as the
    Fill arm of `Pipeline::render_rec` sets them up (src/scene.rs:407-430).  Returns (jobs, device paths, W, H, input bytes)."""
    x0, y0, x1, y1 = sc.view
    lx, ly = math.floor(x0), math.floor(y0)
    W, H = math.ceil(x1) - lx, math.ceil(y1) - ly
    jobs, keep, in_bytes = [], [], 0
    for f in sc.fills:
        bx0, by0 = f.bbox[0], f.bbox[1]
        col_min, row_min, ww, wh = fill_window(f.bbox, lx, ly, W, H)
        tr = Transform.new_translate(-math.floor(bx0), -math.floor(by0)) * Transform.from_array(f.tr)
        dp = rast.upload(f.path)
        keep.append(dp)
        in_bytes += f.path.input_bytes()
        jobs.append(Job(dp, tr, f.fill_rule, ffi.JOB_FILL, layer_ptr, ww, wh, W, origin=row_min * W + col_min, paint=f.paint,
                        path_bbox=f.path_bbox))
    return jobs, keep, W, H, in_bytes


def fixture_fills_host(sc):
    """The same Fill nodes as `fixture_jobs`, as host-side tuples for `GpuRasterizer.render_scene_host`.
    Returns (fills, W, H)."""
    x0, y0, x1, y1 = sc.view
    lx, ly = math.floor(x0), math.floor(y0)
    W, H = math.ceil(x1) - lx, math.ceil(y1) - ly
    fills = []
    for f in sc.fills:
        bx0, by0 = f.bbox[0], f.bbox[1]
        col_min, row_min, ww, wh = fill_window(f.bbox, lx, ly, W, H)
        tr = Transform.new_translate(-math.floor(bx0), -math.floor(by0)) * Transform.from_array(f.tr)
        fills.append((f.path, tr, f.fill_rule, f.paint, f.path_bbox, col_min, row_min, ww, wh))
    return fills, W, H
