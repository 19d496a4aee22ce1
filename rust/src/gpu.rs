//! `GpuRasterizer`: the B200 implementation of `Rasterizer` (src/rasterize.rs:44-101) over the C ABI of
//! `include/rasterize_b200.h`.  `Path::fill`, `Path::mask` and `Scene::render` take `&dyn Rasterizer`, so this is a
//! drop-in: nothing in `Path`, `Curve`, `Transform`, `FillRule` or the trait changes.
//!
//! Uncompiled in the repository that ships this file (no Rust toolchain there); `tests/test_rust_ffi.py` checks the
//! `extern "C"` block against the header.
use crate::{
    rasterize::fill_with_paint, Align, FillRule, ImageMut, LinColor, LineCap, LineJoin, Paint, Path, Pixel, Point, Rasterizer, Scalar, Segment, Size,
    StrokeStyle, Transform,
};
use std::{
    ffi::CStr,
    os::raw::{c_char, c_int, c_void},
    sync::Mutex,
};

// ---- plain-data mirrors of the header's structs -----------------------------------------------------------------
#[repr(C)]
pub struct RgpuPath {
    points: *const f64,
    kinds: *const u8,
    subpath_offsets: *const u32,
    closed: *const u8,
    n_points: u32,
    n_segments: u32,
    n_subpaths: u32,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct RgpuShape {
    start: usize,
    width: usize,
    height: usize,
    row_stride: usize,
    col_stride: usize,
}
#[repr(C)]
pub struct RgpuPixel {
    x: usize,
    y: usize,
    alpha: f64,
}
#[repr(C)]
pub struct RgpuPaint {
    pub kind: i32,
    pub units: i32,
    pub linear_colors: i32,
    pub spread: i32,
    pub tr: [f64; 6],
    pub p0: [f64; 2],
    pub p1: [f64; 2],
    pub r0: f64,
    pub r1: f64,
    pub solid: [f32; 4],
    pub n_stops: u32,
    pub stop_pos: *const f64,
    pub stop_colors: *const f32,
}
#[repr(C)]
pub struct RgpuJob {
    path: *const RgpuDpath,
    tr: [f64; 6],
    fill_rule: i32,
    mode: i32,
    paint: *const RgpuPaint,
    path_bbox: *const f64,
    canvas: *mut c_void,
    origin: usize,
    row_stride: usize,
    width: u32,
    height: u32,
}
#[repr(C)]
pub struct RgpuSceneFill {
    path: *const RgpuPath,
    tr: [f64; 6],
    fill_rule: i32,
    paint: *const RgpuPaint,
    path_bbox: *const f64,
    x: u32,
    y: u32,
    width: u32,
    height: u32,
}
#[repr(C)]
pub struct RgpuStrokeStyle {
    width: f64,
    miter_limit: f64,
    line_join: i32,
    line_cap: i32,
}
/// `rgpu_parse_info`: per-path result of the batch parser
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RgpuParseInfo {
    pub bbox: [f64; 4],
    pub fit_tr: [f64; 6],
    pub fit_width: u32,
    pub fit_height: u32,
    pub n_points: u32,
    pub n_segments: u32,
    pub n_subpaths: u32,
    pub status: i32,
    pub error_offset: u32,
    pub has_bbox: i32,
    pub n_curves: u32,
    pub reserved: u32,
}
#[repr(C)]
pub struct RgpuParseOptions {
    fit_width: u32,
    fit_height: u32,
    fit_align: i32,
}
pub enum RgpuDpathBatch {}
pub enum RgpuCtx {}
pub enum RgpuDpath {}
pub enum RgpuMulti {}

pub const RGPU_OUT_LINCOLOR: c_int = 0;
pub const RGPU_OUT_RGBA8: c_int = 1;
pub const RGPU_OUT_COVERAGE: c_int = 2;

extern "C" {
    fn rgpu_create(device: c_int, flatness: f64, out: *mut *mut RgpuCtx) -> c_int;
    fn rgpu_destroy(ctx: *mut RgpuCtx);
    fn rgpu_last_error(ctx: *const RgpuCtx) -> *const c_char;
    fn rgpu_device_count() -> c_int;
    fn rgpu_flatten(ctx: *mut RgpuCtx, path: *const RgpuPath, tr: *const f64, close: c_int, lines_out: *mut f64, cap: usize, n_out: *mut usize) -> c_int;
    fn rgpu_mask(ctx: *mut RgpuCtx, path: *const RgpuPath, tr: *const f64, fill_rule: c_int, img: *mut f64, shape: RgpuShape) -> c_int;
    fn rgpu_mask_iter(ctx: *mut RgpuCtx, path: *const RgpuPath, tr: *const f64, width: usize, height: usize, fill_rule: c_int, out: *mut RgpuPixel, cap: usize, n_out: *mut usize) -> c_int;
    fn rgpu_fill(ctx: *mut RgpuCtx, path: *const RgpuPath, tr: *const f64, fill_rule: c_int, paint: *const RgpuPaint, path_bbox: *const f64, img: *mut f32, shape: RgpuShape) -> c_int;
    fn rgpu_path_free(ctx: *mut RgpuCtx, p: *mut RgpuDpath);
    fn rgpu_path_stroke(ctx: *mut RgpuCtx, path: *const RgpuPath, style: *const RgpuStrokeStyle, out: *mut *mut RgpuDpath) -> c_int;
    fn rgpu_dpath_stroke(ctx: *mut RgpuCtx, src: *const RgpuDpath, style: *const RgpuStrokeStyle, out: *mut *mut RgpuDpath) -> c_int;
    fn rgpu_dpath_info(p: *const RgpuDpath, n_points: *mut u32, n_segments: *mut u32, n_subpaths: *mut u32) -> c_int;
    fn rgpu_dpath_download(ctx: *mut RgpuCtx, p: *const RgpuDpath, points: *mut f64, kinds: *mut u8, subpath_offsets: *mut u32, closed: *mut u8) -> c_int;
    fn rgpu_parse_svg_batch(ctx: *mut RgpuCtx, text: *const c_char, text_offsets: *const u32, n_paths: usize, opt: *const RgpuParseOptions, out: *mut *mut RgpuDpathBatch, info: *mut RgpuParseInfo) -> c_int;
    fn rgpu_path_batch_info(batch: *const RgpuDpathBatch, n_paths: *mut usize, n_points: *mut u32, n_segments: *mut u32, n_subpaths: *mut u32) -> c_int;
    fn rgpu_path_batch_download(ctx: *mut RgpuCtx, batch: *const RgpuDpathBatch, points: *mut f64, kinds: *mut u8, subpath_offsets: *mut u32, closed: *mut u8, path_subpath_offsets: *mut u32) -> c_int;
    fn rgpu_path_batch_free(ctx: *mut RgpuCtx, batch: *mut RgpuDpathBatch);
    fn rgpu_render_scene_host(ctx: *mut RgpuCtx, fills: *const RgpuSceneFill, n_fills: usize, width: usize, height: usize, bg: *const f32, lin_out: *mut f32, rgba_out: *mut u8) -> c_int;
    fn rgpu_fill_batch_host(ctx: *mut RgpuCtx, all: *const RgpuPath, path_subpath_offsets: *const u32, n_paths: usize, trs: *const f64, fill_rule: c_int, paint: *const RgpuPaint, width: u32, height: u32, out_format: c_int, out_host: *mut c_void) -> c_int;
    fn rgpu_multi_create(devices: *const c_int, n_devices: c_int, flatness: f64, out: *mut *mut RgpuMulti) -> c_int;
    fn rgpu_multi_destroy(m: *mut RgpuMulti);
    fn rgpu_multi_last_error(m: *const RgpuMulti) -> *const c_char;
    fn rgpu_multi_fill_batch_host(m: *mut RgpuMulti, all: *const RgpuPath, path_subpath_offsets: *const u32, n_paths: usize, trs: *const f64, fill_rule: c_int, paint: *const RgpuPaint, width: u32, height: u32, out_format: c_int, out_host: *mut c_void) -> c_int;
    fn rgpu_multi_mask_banded_host(m: *mut RgpuMulti, path: *const RgpuPath, tr: *const f64, fill_rule: c_int, img: *mut c_void, elem_size: usize, width: usize, height: usize, n_bands: u32) -> c_int;
}

/// Flat description of a paint the device can evaluate (returned by `Paint::gpu_desc`, see paint_gpu_desc.patch).
/// Stop colours are in the space the gradient STORES them (after `convert_to_srgb` when `linear_colors` is false).
pub struct PaintDesc {
    pub kind: i32, // 0 solid, 1 linear, 2 radial
    pub units: i32,
    pub linear_colors: bool,
    pub spread: i32,
    pub tr: [f64; 6],
    pub p0: [f64; 2],
    pub p1: [f64; 2],
    pub r0: f64,
    pub r1: f64,
    pub solid: [f32; 4],
    pub stop_pos: Vec<f64>,
    pub stop_colors: Vec<f32>,
}

impl PaintDesc {
    pub fn solid(c: LinColor) -> Self {
        let c: [f32; 4] = c.into();
        Self { kind: 0, units: 0, linear_colors: true, spread: 0, tr: [1.0, 0.0, 0.0, 0.0, 1.0, 0.0], p0: [0.0; 2], p1: [0.0; 2], r0: 0.0, r1: 0.0,
               solid: c, stop_pos: Vec::new(), stop_colors: Vec::new() }
    }
    fn ffi(&self) -> RgpuPaint {
        RgpuPaint { kind: self.kind, units: self.units, linear_colors: self.linear_colors as i32, spread: self.spread, tr: self.tr, p0: self.p0,
                    p1: self.p1, r0: self.r0, r1: self.r1, solid: self.solid, n_stops: self.stop_pos.len() as u32,
                    stop_pos: self.stop_pos.as_ptr(), stop_colors: self.stop_colors.as_ptr() }
    }
}

/// Flat re-encoding of `Path` (src/path.rs:227-233): `Segment` is a Rust enum and cannot cross the FFI.
/// `append` adds one more path to the same arrays (batches: `offsets_per_path` marks where each path's subpaths end).
#[derive(Default)]
struct FlatPath {
    points: Vec<f64>,
    kinds: Vec<u8>,
    offsets: Vec<u32>,
    closed: Vec<u8>,
    path_subpath_offsets: Vec<u32>,
}

impl FlatPath {
    fn new(path: &Path) -> Self {
        let mut flat = Self::default();
        flat.append(path);
        flat
    }
    fn append(&mut self, path: &Path) {
        if self.offsets.is_empty() {
            self.offsets.push(0);
            self.path_subpath_offsets.push(0);
        }
        for sp in path.subpaths() {
            // src/path.rs:329-334
            for seg in sp.segments() {
                // src/path.rs:175-177
                let pts: &[Point] = match seg {
                    Segment::Line(l) => &l.0,
                    Segment::Quad(q) => &q.0,
                    Segment::Cubic(c) => &c.0,
                };
                self.kinds.push(pts.len() as u8);
                for p in pts {
                    self.points.push(p.x());
                    self.points.push(p.y());
                }
            }
            self.offsets.push(self.kinds.len() as u32);
            self.closed.push(sp.is_closed() as u8);
        }
        self.path_subpath_offsets.push(self.closed.len() as u32);
    }
    fn ffi(&self) -> RgpuPath {
        RgpuPath {
            points: self.points.as_ptr(),
            kinds: self.kinds.as_ptr(),
            subpath_offsets: if self.closed.is_empty() { std::ptr::null() } else { self.offsets.as_ptr() },
            closed: self.closed.as_ptr(),
            n_points: (self.points.len() / 2) as u32,
            n_segments: self.kinds.len() as u32,
            n_subpaths: self.closed.len() as u32,
        }
    }
}

fn shape_of<I: ImageMut + ?Sized>(img: &I) -> RgpuShape {
    let s = img.shape(); // src/image.rs:6-17
    RgpuShape { start: s.start, width: s.width, height: s.height, row_stride: s.row_stride, col_stride: s.col_stride }
}

fn rule(fill_rule: FillRule) -> c_int {
    match fill_rule {
        FillRule::NonZero => 0,
        FillRule::EvenOdd => 1,
    }
}

/// One CUDA context per rasterizer.  The C context is single-threaded, hence the `Mutex` (`Rasterizer` methods take `&self`).
pub struct GpuRasterizer {
    ctx: Mutex<*mut RgpuCtx>,
}
unsafe impl Send for GpuRasterizer {}
unsafe impl Sync for GpuRasterizer {}

impl GpuRasterizer {
    pub fn new(device: i32, flatness: Scalar) -> Self {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { rgpu_create(device, flatness, &mut ctx) };
        if rc != 0 {
            panic!("rasterize_b200: {}", unsafe { CStr::from_ptr(rgpu_last_error(std::ptr::null())) }.to_string_lossy());
        }
        Self { ctx: Mutex::new(ctx) }
    }

    pub fn device_count() -> usize {
        unsafe { rgpu_device_count() as usize }
    }

    /// The reference panics on bad input ("cannot flatten segment with NaN", src/path.rs:765-767); so does the shim.
    fn check(ctx: *mut RgpuCtx, rc: c_int) {
        if rc != 0 {
            panic!("rasterize_b200: {}", unsafe { CStr::from_ptr(rgpu_last_error(ctx)) }.to_string_lossy());
        }
    }

    /// `Path::flatten(tr, flatness, close)` on the device: the same lines, in the same order, bit for bit.
    pub fn flatten(&self, path: &Path, tr: Transform, close: bool) -> Vec<crate::Line> {
        let flat = FlatPath::new(path);
        let trm: [Scalar; 6] = tr.into();
        let ctx = self.ctx.lock().unwrap();
        let mut n = 0usize;
        let mut buf: Vec<f64> = Vec::with_capacity(4 * (flat.kinds.len() * 32 + 64));
        loop {
            let cap = buf.capacity() / 4;
            let rc = unsafe { rgpu_flatten(*ctx, &flat.ffi(), trm.as_ptr(), close as c_int, buf.as_mut_ptr(), cap, &mut n) };
            if rc == -5 && n > cap {
                buf.reserve(4 * n); // RGPU_ERR_CAPACITY: n is the full count
                continue;
            }
            Self::check(*ctx, rc);
            break;
        }
        unsafe { buf.set_len(4 * n) };
        buf.chunks_exact(4).map(|l| crate::Line::new((l[0], l[1]), (l[2], l[3]))).collect()
    }

    /// `Path::stroke(style)` (src/path.rs:374-415) on the device.  The C ABI hands the outline back device-resident
    /// (`rgpu_dpath`, for callers that rasterize it next); this method downloads it into an ordinary `Path`.
    pub fn stroke(&self, path: &Path, style: StrokeStyle) -> Path {
        let flat = FlatPath::new(path);
        let (line_join, miter_limit) = match style.line_join {
            LineJoin::Miter(limit) => (0, limit),
            LineJoin::Bevel => (1, 4.0),
            LineJoin::Round => (2, 4.0),
        };
        let line_cap = match style.line_cap {
            LineCap::Butt => 0,
            LineCap::Square => 1,
            LineCap::Round => 2,
        };
        let st = RgpuStrokeStyle { width: style.width, miter_limit, line_join, line_cap };
        let ctx = self.ctx.lock().unwrap();
        let mut dp = std::ptr::null_mut();
        Self::check(*ctx, unsafe { rgpu_path_stroke(*ctx, &flat.ffi(), &st, &mut dp) });
        let (mut n_pts, mut n_seg, mut n_sub) = (0u32, 0u32, 0u32);
        unsafe { rgpu_dpath_info(dp, &mut n_pts, &mut n_seg, &mut n_sub) };
        let mut points = vec![0.0f64; 2 * n_pts as usize];
        let mut kinds = vec![0u8; n_seg as usize];
        let mut offsets = vec![0u32; n_sub as usize + 1];
        let mut closed = vec![0u8; n_sub as usize];
        let rc = unsafe { rgpu_dpath_download(*ctx, dp, points.as_mut_ptr(), kinds.as_mut_ptr(), offsets.as_mut_ptr(), closed.as_mut_ptr()) };
        unsafe { rgpu_path_free(*ctx, dp) };
        Self::check(*ctx, rc);
        // back into `Path { segments, subpaths, closed }` (src/path.rs:227-240)
        let mut segments = Vec::with_capacity(kinds.len());
        let mut at = 0usize;
        let pt = |i: usize| Point::new(points[2 * i], points[2 * i + 1]);
        for k in &kinds {
            segments.push(match k {
                2 => Segment::Line(crate::Line::new(pt(at), pt(at + 1))),
                3 => Segment::Quad(crate::Quad::new(pt(at), pt(at + 1), pt(at + 2))),
                _ => Segment::Cubic(crate::Cubic::new(pt(at), pt(at + 1), pt(at + 2), pt(at + 3))),
            });
            at += *k as usize;
        }
        if segments.is_empty() {
            return Path::empty();
        }
        Path::new(segments, offsets.iter().map(|o| *o as usize).collect(), closed.iter().map(|c| *c != 0).collect())
    }

    /// `str::parse::<Path>()` for a batch of SVG path strings, with `Path::bbox(identity)` and (when `fit` is given)
    /// `fit_size(bbox, size, align)` per path, in one device call (src/svg.rs:241-421, src/path.rs:428-451,
    /// src/geometry.rs:490-516).  A string that does not parse comes back as `Err((kind, byte offset))`.
    pub fn parse_svg_batch(&self, strings: &[&str], fit: Option<(Size, Align)>) -> Vec<Result<(Path, RgpuParseInfo), (i32, u32)>> {
        let mut text = String::new();
        let mut off = vec![0u32];
        for s in strings {
            text.push_str(s);
            off.push(text.len() as u32);
        }
        let opt = match fit {
            Some((size, align)) => RgpuParseOptions { fit_width: size.width as u32, fit_height: size.height as u32,
                                                      fit_align: match align { Align::Min => 0, Align::Mid => 1, Align::Max => 2 } },
            None => RgpuParseOptions { fit_width: 0, fit_height: 0, fit_align: -1 },
        };
        let mut info = vec![RgpuParseInfo::default(); strings.len()];
        let ctx = self.ctx.lock().unwrap();
        let mut b = std::ptr::null_mut();
        Self::check(*ctx, unsafe { rgpu_parse_svg_batch(*ctx, text.as_ptr() as *const c_char, off.as_ptr(), strings.len(), &opt, &mut b, info.as_mut_ptr()) });
        let (mut n, mut n_pts, mut n_seg, mut n_sub) = (0usize, 0u32, 0u32, 0u32);
        unsafe { rgpu_path_batch_info(b, &mut n, &mut n_pts, &mut n_seg, &mut n_sub) };
        let mut points = vec![0.0f64; 2 * n_pts as usize];
        let mut kinds = vec![0u8; n_seg as usize];
        let mut sp = vec![0u32; n_sub as usize + 1];
        let mut closed = vec![0u8; n_sub as usize];
        let mut psp = vec![0u32; n + 1];
        let rc = unsafe { rgpu_path_batch_download(*ctx, b, points.as_mut_ptr(), kinds.as_mut_ptr(), sp.as_mut_ptr(), closed.as_mut_ptr(), psp.as_mut_ptr()) };
        unsafe { rgpu_path_batch_free(*ctx, b) };
        Self::check(*ctx, rc);
        let pt = |i: usize| Point::new(points[2 * i], points[2 * i + 1]);
        let mut at = 0usize;
        (0..n)
            .map(|i| {
                if info[i].status != 0 {
                    return Err((info[i].status, info[i].error_offset));
                }
                let (s0, s1) = (psp[i] as usize, psp[i + 1] as usize);
                if s0 == s1 {
                    return Ok((Path::empty(), info[i]));
                }
                let (k0, k1) = (sp[s0] as usize, sp[s1] as usize);
                let mut segments = Vec::with_capacity(k1 - k0);
                for k in &kinds[k0..k1] {
                    segments.push(match k {
                        2 => Segment::Line(crate::Line::new(pt(at), pt(at + 1))),
                        3 => Segment::Quad(crate::Quad::new(pt(at), pt(at + 1), pt(at + 2))),
                        _ => Segment::Cubic(crate::Cubic::new(pt(at), pt(at + 1), pt(at + 2), pt(at + 3))),
                    });
                    at += *k as usize;
                }
                let subpaths = sp[s0..=s1].iter().map(|o| *o as usize - k0).collect();
                let closed = closed[s0..s1].iter().map(|c| *c != 0).collect();
                Ok((Path::new(segments, subpaths, closed), info[i]))
            })
            .collect()
    }

    /// `ImageOwned::new_default(size)` + `Path::fill` for a batch of independent paths (glyph batches): the images come back
    /// back to back, `paths.len() * size.height * size.width` LinColor pixels.
    pub fn fill_batch(&self, paths: &[Path], tr: Transform, fill_rule: FillRule, paint: &PaintDesc, size: Size) -> Vec<LinColor> {
        let mut flat = FlatPath::default();
        paths.iter().for_each(|p| flat.append(p));
        let trm: [Scalar; 6] = tr.into();
        let trs: Vec<f64> = std::iter::repeat(trm).take(paths.len()).flatten().collect();
        let mut out = vec![LinColor::default(); paths.len() * size.width * size.height];
        let ctx = self.ctx.lock().unwrap();
        Self::check(*ctx, unsafe {
            rgpu_fill_batch_host(*ctx, &flat.ffi(), flat.path_subpath_offsets.as_ptr(), paths.len(), trs.as_ptr(), rule(fill_rule), &paint.ffi(),
                                 size.width as u32, size.height as u32, RGPU_OUT_LINCOLOR, out.as_mut_ptr() as *mut c_void)
        });
        out
    }
}

impl Drop for GpuRasterizer {
    fn drop(&mut self) {
        unsafe { rgpu_destroy(*self.ctx.lock().unwrap()) }
    }
}

impl Rasterizer for GpuRasterizer {
    fn name(&self) -> &str {
        "gpu-signed-difference"
    }

    fn mask(&self, path: &Path, tr: Transform, img: &mut dyn ImageMut<Pixel = Scalar>, fill_rule: FillRule) {
        let flat = FlatPath::new(path);
        let shape = shape_of(img);
        let trm: [Scalar; 6] = tr.into(); // [m00, m01, m02, m10, m11, m12], see paint_gpu_desc.patch
        let ctx = self.ctx.lock().unwrap();
        Self::check(*ctx, unsafe { rgpu_mask(*ctx, &flat.ffi(), trm.as_ptr(), rule(fill_rule), img.data_mut().as_mut_ptr(), shape) });
    }

    fn mask_iter(&self, path: &Path, tr: Transform, size: Size, fill_rule: FillRule) -> Box<dyn Iterator<Item = Pixel> + '_> {
        let flat = FlatPath::new(path);
        // the list is compacted on the device: only the yielded pixels come down.  First guess: an outline's worth of pixels;
        // RGPU_ERR_CAPACITY returns the full count and the call is repeated once with room for it
        let mut buf: Vec<RgpuPixel> = Vec::with_capacity((size.width * size.height).min(1 << 20));
        let mut n = 0usize;
        let trm: [Scalar; 6] = tr.into();
        let ctx = self.ctx.lock().unwrap();
        loop {
            let cap = buf.capacity();
            let rc = unsafe {
                rgpu_mask_iter(*ctx, &flat.ffi(), trm.as_ptr(), size.width, size.height, rule(fill_rule), buf.as_mut_ptr(), cap, &mut n)
            };
            if rc == -5 && n > cap {
                buf.reserve(n);
                continue;
            }
            Self::check(*ctx, rc);
            break;
        }
        unsafe { buf.set_len(n) };
        Box::new(buf.into_iter().map(|p| Pixel { x: p.x, y: p.y, alpha: p.alpha }))
    }

    /// Overrides the default `fill` (src/rasterize.rs:70-100) when the paint can describe itself; an opaque `&dyn Paint`
    /// takes the default body (`fill_with_paint`, factored out by paint_gpu_desc.patch) over this rasterizer's `mask_iter`.
    fn fill(&self, path: &Path, tr: Transform, fill_rule: FillRule, paint: &dyn Paint, img: &mut dyn ImageMut<Pixel = LinColor>) {
        let Some(desc) = paint.gpu_desc() else {
            let pixels = self.mask_iter(path, tr, img.size(), fill_rule);
            return fill_with_paint(pixels, path, tr, paint, img);
        };
        let flat = FlatPath::new(path);
        let bbox = path.bbox(Transform::identity()).map(|b| [b.min().x(), b.min().y(), b.max().x(), b.max().y()]);
        let shape = shape_of(img);
        let trm: [Scalar; 6] = tr.into();
        let ctx = self.ctx.lock().unwrap();
        Self::check(*ctx, unsafe {
            rgpu_fill(*ctx, &flat.ffi(), trm.as_ptr(), rule(fill_rule), &desc.ffi(), bbox.as_ref().map_or(std::ptr::null(), |b| b.as_ptr()),
                      img.data_mut().as_mut_ptr() as *mut f32, shape)
        });
    }
}

/// Every GPU of the box behind one object: batches shard by path, one huge canvas by scanline bands; no
/// collective, every device copies its shard into its own region of the caller's buffer.
pub struct MultiGpuRasterizer {
    m: Mutex<*mut RgpuMulti>,
}
unsafe impl Send for MultiGpuRasterizer {}
unsafe impl Sync for MultiGpuRasterizer {}

impl MultiGpuRasterizer {
    pub fn new(devices: &[i32], flatness: Scalar) -> Self {
        let mut m = std::ptr::null_mut();
        let rc = unsafe { rgpu_multi_create(devices.as_ptr(), devices.len() as c_int, flatness, &mut m) };
        if rc != 0 {
            panic!("rasterize_b200: {}", unsafe { CStr::from_ptr(rgpu_multi_last_error(std::ptr::null())) }.to_string_lossy());
        }
        Self { m: Mutex::new(m) }
    }
    fn check(m: *mut RgpuMulti, rc: c_int) {
        if rc != 0 {
            panic!("rasterize_b200: {}", unsafe { CStr::from_ptr(rgpu_multi_last_error(m)) }.to_string_lossy());
        }
    }
    /// `GpuRasterizer::fill_batch` over all devices.
    pub fn fill_batch(&self, paths: &[Path], tr: Transform, fill_rule: FillRule, paint: &PaintDesc, size: Size) -> Vec<LinColor> {
        let mut flat = FlatPath::default();
        paths.iter().for_each(|p| flat.append(p));
        let trm: [Scalar; 6] = tr.into();
        let trs: Vec<f64> = std::iter::repeat(trm).take(paths.len()).flatten().collect();
        let mut out = vec![LinColor::default(); paths.len() * size.width * size.height];
        let m = self.m.lock().unwrap();
        Self::check(*m, unsafe {
            rgpu_multi_fill_batch_host(*m, &flat.ffi(), flat.path_subpath_offsets.as_ptr(), paths.len(), trs.as_ptr(), rule(fill_rule), &paint.ffi(),
                                       size.width as u32, size.height as u32, RGPU_OUT_LINCOLOR, out.as_mut_ptr() as *mut c_void)
        });
        out
    }
    /// `Path::mask` on a huge dense canvas, split into scanline bands over the devices (`n_bands` = 0: eight per device).
    pub fn mask_banded(&self, path: &Path, tr: Transform, fill_rule: FillRule, img: &mut [Scalar], size: Size, n_bands: u32) {
        assert_eq!(img.len(), size.width * size.height);
        let flat = FlatPath::new(path);
        let trm: [Scalar; 6] = tr.into();
        let m = self.m.lock().unwrap();
        Self::check(*m, unsafe {
            rgpu_multi_mask_banded_host(*m, &flat.ffi(), trm.as_ptr(), rule(fill_rule), img.as_mut_ptr() as *mut c_void,
                                        std::mem::size_of::<Scalar>(), size.width, size.height, n_bands)
        });
    }
}

impl Drop for MultiGpuRasterizer {
    fn drop(&mut self) {
        unsafe { rgpu_multi_destroy(*self.m.lock().unwrap()) }
    }
}
