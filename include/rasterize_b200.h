/* rasterize_b200 — C ABI of the B200-native fill pipeline (flatten -> signed-difference raster ->
 * paint/composite).  This is the drop-in boundary a Rust `GpuRasterizer: Rasterizer` binds over FFI
 * (see INTEGRATION.md for the `extern "C"` block and build.rs).  Plain pointers and sizes only.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * crate aslpavel/rasterize v0.6.7).  All functions return RGPU_OK (0) or a negative status; the text
 * of the last error is available from rgpu_last_error().  The Rust shim turns a non-zero status
 * into `panic!`, which is the reference's behaviour on bad input (src/path.rs:765-767).
 *
 * There is NO CPU fallback: if no CUDA device is usable rgpu_create fails with RGPU_ERR_CUDA.
 *
 * Threading: a context may be used from one thread at a time (the reference rasterizers are
 * stateless `&self` objects, src/rasterize.rs:279-296; the shim wraps the context in a Mutex).
 */
#ifndef RASTERIZE_B200_H
#define RASTERIZE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGPU_OK 0
#define RGPU_ERR_INVALID (-1)   /* bad argument (null pointer, flatness <= 0, bad kind ...) */
#define RGPU_ERR_CUDA (-2)      /* CUDA runtime error, text in rgpu_last_error */
#define RGPU_ERR_NAN (-3)       /* a control point is NaN: reference panics "cannot flatten segment with NaN" */
#define RGPU_ERR_DEPTH (-4)     /* subdivision deeper than the device stack (reference would recurse without bound) */
#define RGPU_ERR_CAPACITY (-5)  /* output buffer supplied by the caller is too small */
#define RGPU_ERR_WINDING (-6)   /* a non-zero winding number reached the guard of the 32-bit fixed-point cells (asynchronous
                                   submissions only: the *_sync and host-buffer entry points re-run such a batch themselves) */

typedef struct rgpu_ctx rgpu_ctx;
typedef struct rgpu_dpath rgpu_dpath; /* device-resident path */

/* FillRule, src/path.rs:21-29 */
enum { RGPU_NONZERO = 0, RGPU_EVENODD = 1 };

/* Flat re-encoding of `Path { segments, subpaths, closed }` (src/path.rs:227-233; `Segment` is a Rust
 * enum and not FFI-stable, src/curve.rs:905-909).  Segment i owns `kinds[i]` consecutive points
 * (2 = Line, 3 = Quad, 4 = Cubic; src/curve.rs:163,349,614), segments are stored back to back in `points`.
 * Subpath s is segments [subpath_offsets[s], subpath_offsets[s+1]). */
typedef struct {
    const double* points;            /* 2 * n_points, (x, y) pairs */
    const uint8_t* kinds;            /* n_segments */
    const uint32_t* subpath_offsets; /* n_subpaths + 1 (may be NULL when n_subpaths == 0) */
    const uint8_t* closed;           /* n_subpaths */
    uint32_t n_points, n_segments, n_subpaths;
} rgpu_path;

/* `Shape`, src/image.rs:6-17 (element strides, not bytes) */
typedef struct {
    size_t start, width, height, row_stride, col_stride;
} rgpu_shape;

/* `Pixel`, src/rasterize.rs (x, y, alpha) as yielded by `Rasterizer::mask_iter` */
typedef struct {
    size_t x, y;
    double alpha;
} rgpu_pixel;

/* Paint description: `impl Paint for LinColor` (src/color.rs:357-374), `GradLinear` (src/grad.rs:150-160),
 * `GradRadial` (src/grad.rs:307-317).  Stop colours are in the space the gradient STORES them, i.e. after
 * `convert_to_srgb` when `linear_colors == 0` (src/grad.rs:165-191). */
enum { RGPU_PAINT_SOLID = 0, RGPU_PAINT_LINEAR = 1, RGPU_PAINT_RADIAL = 2 };
enum { RGPU_UNITS_USER_SPACE = 0, RGPU_UNITS_BOUNDING_BOX = 1 }; /* `Units`, src/rasterize.rs:171-176 */
enum { RGPU_SPREAD_PAD = 0, RGPU_SPREAD_REPEAT = 1, RGPU_SPREAD_REFLECT = 2 }; /* `GradSpread`, src/grad.rs:13-20 */
#define RGPU_MAX_STOPS 32
typedef struct {
    int32_t kind, units, linear_colors, spread;
    double tr[6];    /* `Paint::transform` */
    double p0[2];    /* linear: start ; radial: center */
    double p1[2];    /* linear: end   ; radial: fcenter */
    double r0, r1;   /* radial: radius, fradius */
    float solid[4];  /* solid: premultiplied linear RGBA */
    uint32_t n_stops;
    const double* stop_pos;   /* n_stops */
    const float* stop_colors; /* 4 * n_stops, premultiplied, stored space */
} rgpu_paint;

/* ---- context ------------------------------------------------------------------------------------ */
/* `SignedDifferenceRasterizer::new(flatness)` (src/rasterize.rs:284-296) on CUDA device `device`. */
int rgpu_create(int device, double flatness, rgpu_ctx** out);
void rgpu_destroy(rgpu_ctx* ctx);
/* `Rasterizer::name` (src/rasterize.rs:46, 357-359) */
const char* rgpu_name(void);
const char* rgpu_last_error(const rgpu_ctx* ctx); /* ctx may be NULL: error of the last failed rgpu_create */
int rgpu_device_count(void);

/* ---- trait-level entry points: HOST buffers in, HOST buffers out --------------------------------- */
/* `Path::flatten(tr, flatness, close)` (src/path.rs:418-425, 744-795): lines_out receives 4 doubles per line
 * (x0,y0,x1,y1) in the reference's order; *n_out is always the full count (RGPU_ERR_CAPACITY if > cap). */
int rgpu_flatten(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int close, double* lines_out, size_t cap,
                 size_t* n_out);
/* `Rasterizer::mask` (src/rasterize.rs:50-56, 299-311): img is a strided f64 host image, assumed zero on
 * entry (src/path.rs:511-512) and overwritten with coverage; its last column is the anti-alias overflow
 * column exactly as in the reference (src/rasterize.rs:372). */
int rgpu_mask(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, double* img, rgpu_shape shape);
/* Same semantics, dense row-major f32 host image (the device-native format). */
int rgpu_mask_f32(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, float* img, size_t width,
                  size_t height);
/* `Rasterizer::mask_iter` (src/rasterize.rs:61-67, 313-355): pixels with abs(alpha) >= 1e-6 in row-major
 * order; internal (width+1) canvas, overflow column dropped.  The list is compacted on the device; the first min(n, cap)
 * records are written to `out` (NULL: none).  *n_out is the full count; RGPU_ERR_CAPACITY when it exceeds cap. */
int rgpu_mask_iter(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], size_t width, size_t height, int fill_rule,
                   rgpu_pixel* out, size_t cap, size_t* n_out);
/* Dense form of mask_iter: coverage[y*width + x] (0 where the iterator yields nothing). */
int rgpu_coverage_f32(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, float* out, size_t width,
                      size_t height);
/* Default `Rasterizer::fill` (src/rasterize.rs:70-115): blend `paint` through the coverage of `path` over the
 * strided LinColor (f32 x 4, premultiplied linear) host image.  path_bbox = `path.bbox(identity)` as
 * (minx,miny,maxx,maxy), needed only for RGPU_UNITS_BOUNDING_BOX (NULL there => no-op, src/rasterize.rs:87-90).
 * A singular paint transform is a silent no-op (src/rasterize.rs:92). */
int rgpu_fill(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, const rgpu_paint* paint,
              const double* path_bbox, float* img, rgpu_shape shape);

/* One Fill node of a host-side scene: what the Fill arm of `Pipeline::render_rec` holds per node (src/scene.rs:407-430).
 * (x, y, width, height) is the `view_mut` window of the layer the node draws into, `tr` = align * node transform. */
typedef struct {
    const rgpu_path* path;
    double tr[6];
    int32_t fill_rule;
    const rgpu_paint* paint;
    const double* path_bbox; /* bounding-box units only, else NULL */
    uint32_t x, y, width, height;
} rgpu_scene_fill;
/* `Scene::render` of a Fill-only pipeline (src/scene.rs:186-199, 384-435) followed by the export every CLI run does
 * (`ImageOwned<LinColor>` -> RGBA8, src/color.rs:164-175), HOST buffers in and out: a width x height layer is created with
 * `bg` (NULL = transparent), the fills are blended in order by the scene compositor (one raster launch), and the image is
 * returned as RGBA8 (rgba_out, 4 B per pixel over PCIe) and / or LinColor (lin_out, f32 x 4); either may be NULL. */
int rgpu_render_scene_host(rgpu_ctx* ctx, const rgpu_scene_fill* fills, size_t n_fills, size_t width, size_t height, const float* bg,
                           float* lin_out, uint8_t* rgba_out);

/* ---- device-resident entry points (inputs and outputs stay in HBM) -------------------------------- */
int rgpu_path_upload(rgpu_ctx* ctx, const rgpu_path* path, rgpu_dpath** out);
void rgpu_path_free(rgpu_ctx* ctx, rgpu_dpath* p);

/* `StrokeStyle` (src/path.rs:121-136), `LineJoin` (:78-93, the miter limit lives in `Miter(limit)`, default 4.0) and
 * `LineCap` (:103-110). */
enum { RGPU_JOIN_MITER = 0, RGPU_JOIN_BEVEL = 1, RGPU_JOIN_ROUND = 2 };
enum { RGPU_CAP_BUTT = 0, RGPU_CAP_SQUARE = 1, RGPU_CAP_ROUND = 2 };
typedef struct {
    double width;
    double miter_limit; /* used by RGPU_JOIN_MITER only */
    int32_t line_join, line_cap;
} rgpu_stroke_style;
/* `Path::stroke` (src/path.rs:374-415; `stroke_segment` :692-706, `stroke_close` :708-732; curve offsets, joins and caps
 * src/curve.rs:978-1078, 1283-1433) computed on the device: the host path goes up once, the outline of the stroke comes
 * back as a device-resident path that every entry point taking an `rgpu_dpath` accepts (config 5's pre-step without a
 * round trip).  One thread per (segment, direction) unit of the reference's walk, three passes (pieces, joins, emit), the
 * segments in the reference's order.  With Miter / Bevel joins and Butt / Square caps the segment list is bit-identical to
 * the reference's; round joins and caps go through sin / cos / tan / acos and agree to a few ulp.
 * Deviations: a cubic whose four control points coincide makes the reference panic (`ends`, src/curve.rs:660-666) and is
 * treated here as having the tangent of its first two points; NaN input that would spin the reference's arc iterator for
 * ever yields no arc. */
int rgpu_path_stroke(rgpu_ctx* ctx, const rgpu_path* path, const rgpu_stroke_style* style, rgpu_dpath** out);
/* The same for a path that is already on the device (an uploaded path, an element of an uploaded or parsed batch): its
 * control points are read where they lie; only its item list (8 B per segment) visits the host, to lay out the unit table. */
int rgpu_dpath_stroke(rgpu_ctx* ctx, const rgpu_dpath* src, const rgpu_stroke_style* style, rgpu_dpath** out);
/* Size of a device path made by rgpu_path_upload / rgpu_path_stroke, and its download in the `rgpu_path` encoding:
 * points[2 * n_points], kinds[n_segments], subpath_offsets[n_subpaths + 1], closed[n_subpaths]. */
int rgpu_dpath_info(const rgpu_dpath* p, uint32_t* n_points, uint32_t* n_segments, uint32_t* n_subpaths);
int rgpu_dpath_download(rgpu_ctx* ctx, const rgpu_dpath* p, double* points, uint8_t* kinds, uint32_t* subpath_offsets, uint8_t* closed);

/* One fill job of a batch.  `canvas` is a DEVICE pointer; the job covers the `width` x `height` window whose
 * top-left element is canvas[origin] with `row_stride` elements between rows (elements = f32 for masks,
 * 4 x f32 for colour).  This is the device form of `Path::fill` on a `view_mut` sub-image
 * (src/scene.rs:412-429) and of `Path::mask`. */
enum {
    RGPU_JOB_MASK = 0,     /* Rasterizer::mask semantics  -> f32 coverage, last column = overflow column */
    RGPU_JOB_COVERAGE = 1, /* mask_iter semantics         -> f32 coverage, (width+1) internal columns */
    RGPU_JOB_FILL = 2,     /* Rasterizer::fill semantics  -> blend paint over LinColor canvas */
    RGPU_JOB_RENDER = 3    /* `ImageOwned::new_default(size)` + Rasterizer::fill: the job CREATES its LinColor window
                              (transparent, as Layer::new without a background) and fills onto it — bit-identical to
                              RGPU_JOB_FILL on a zeroed window, but every pixel is written once and none is read
                              (16 B per pixel; the glyph-batch form of BASELINE config 4) */
};
typedef struct {
    const rgpu_dpath* path;
    double tr[6];
    int32_t fill_rule;
    int32_t mode;
    const rgpu_paint* paint;  /* RGPU_JOB_FILL only (host pointer, copied at submission) */
    const double* path_bbox;  /* RGPU_JOB_FILL with bounding-box units */
    void* canvas;             /* device pointer */
    size_t origin, row_stride; /* in elements */
    uint32_t width, height;
} rgpu_job;

/* Batch flags */
#define RGPU_BATCH_ORDERED 0u     /* jobs may overlap; composited in submission order (Scene::render Fill arm) */
#define RGPU_BATCH_INDEPENDENT 1u /* caller guarantees disjoint outputs: one raster launch for all jobs */

/* Flatten + bin + rasterize `n_jobs` jobs on the context's stream (asynchronous; call rgpu_sync or
 * rgpu_batch_status afterwards).  Replaces the per-node loop of `Pipeline::render_rec` (src/scene.rs:397-435)
 * for Fill nodes and a batch of `Path::mask` calls. */
int rgpu_render_batch(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags);
/* Waits for the stream and reports device-side errors of the last batch (NaN, depth, internal capacity —
 * the latter is retried transparently by the host-buffer entry points and by rgpu_render_batch_sync). */
int rgpu_batch_status(rgpu_ctx* ctx);
/* rgpu_render_batch + rgpu_batch_status with transparent scratch growth and re-run on internal overflow. */
int rgpu_render_batch_sync(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags);
/* Every Fill of one layer in ONE raster launch: the Fill arm of `Pipeline::render_rec` (src/scene.rs:397-435) for all Fill
 * nodes drawn onto a layer, fused with `Layer::new` (src/scene.rs:483-501) and, optionally, the RGBA8 export
 * (`From<LinColor> for RGBA`, src/color.rs:164-175).  `layer_dev` is a dense width x height LinColor device image; every
 * job must be an RGPU_JOB_FILL whose canvas is `layer_dev` with row_stride == width (origin selects its `view_mut`
 * window).  Jobs are blended in submission order, exactly like RGPU_BATCH_ORDERED, but a layer tile stays on the SM while
 * all its fills are applied, so the layer is written once.
 *   fresh != 0: the layer is created here — every pixel starts from bg (NULL = transparent), as `Layer::new` does;
 *   fresh == 0: the fills blend over what the layer already holds.
 * rgba_dev (may be NULL): width x height RGBA8 device image that receives the finished layer. */
int rgpu_render_scene(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, float* layer_dev, size_t width, size_t height, int fresh,
                      const float* bg, uint8_t* rgba_dev);
/* rgpu_render_scene + rgpu_batch_status with transparent scratch growth and re-run on internal overflow. */
int rgpu_render_scene_sync(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, float* layer_dev, size_t width, size_t height, int fresh,
                           const float* bg, uint8_t* rgba_dev);
/* Winding cells are 32-bit fixed point.  integer_bits = 8 (default): Q7.24, 6e-8 of a pixel; the integer part wraps modulo 256
 * windings, which the even-odd rule cannot see and the non-zero rule sees only for a winding within 1 of a non-zero multiple of
 * 256 (the reference accumulates f64, src/rasterize.rs:478-493).  The row scans report a non-zero winding of 120 or more
 * (RGPU_ERR_WINDING from rgpu_batch_status); every *_sync / host-buffer entry point then re-runs the batch with
 * integer_bits = 14 (Q13.18, 4e-6 of a pixel, windings up to 8191).  This call selects the format up front for the
 * asynchronous entry points.  Not covered: 128 or more coincident same-direction edges through ONE pixel cell (identical
 * stacked copies) can wrap that cell before any winding is seen. */
int rgpu_set_winding_bits(rgpu_ctx* ctx, int integer_bits);
/* Lines produced by the flatten stage of the last completed batch, and kernels launched since create. */
int rgpu_last_counts(rgpu_ctx* ctx, uint64_t* n_lines, uint64_t* n_line_refs, uint64_t* n_launches);

/* Bytes the last rgpu_mask / rgpu_mask_f32 call moved over PCIe (path upload; image download: f32 rows + f64 rows). */
int rgpu_last_transfer_bytes(rgpu_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* Optional per-stage device timing of batches (CUDA events on the context's stream around the flatten, bin and
 * raster stages).  rgpu_last_stage_ms is valid after rgpu_batch_status / *_sync: out = {flatten, bin, raster} ms. */
int rgpu_set_profiling(rgpu_ctx* ctx, int enable);
int rgpu_last_stage_ms(rgpu_ctx* ctx, float out[3]);

/* LinColor (f32x4) device image -> RGBA8 device image: `From<LinColor> for RGBA` (src/color.rs:164-175) with the
 * x86 `l2s` polynomial (src/simd/x86.rs:197-214). n = pixels. */
int rgpu_to_rgba8_dev(rgpu_ctx* ctx, const float* lin_dev, uint8_t* rgba_dev, size_t n_pixels);
/* fill a LinColor device image with a constant (Layer::new with bg, src/scene.rs:483-501) */
int rgpu_fill_color_dev(rgpu_ctx* ctx, float* lin_dev, size_t n_pixels, const float color[4]);

/* `Layer::compose` on device (src/scene.rs:532-565) for the Clip and Opacity arms of `Pipeline::render_rec`
 * (src/scene.rs:436-457).  Layers are dense device images; `*_origin` is the element index of the top-left pixel of the
 * intersection rectangle inside each layer and `*_stride` its row pitch, both in pixels; width x height is that
 * rectangle (computed by the caller as Layer::compose does from the layers' x / y / size).
 *   rgpu_layer_scale_by_mask_dev:  lin[p] = lin[p] * mask[p]                 (Clip: child_layer.compose(mask_layer, ..))
 *   rgpu_layer_blend_over_dev:     dst[p] = dst[p].blend_over(src[p] [* opacity])   (Clip / Opacity: layer.compose(child_layer, ..))
 * `use_opacity == 0` leaves the source unscaled (the Clip arm), otherwise it is multiplied by `opacity` as f32. */
int rgpu_layer_scale_by_mask_dev(rgpu_ctx* ctx, float* lin_dev, size_t lin_origin, size_t lin_stride, const float* mask_dev,
                                 size_t mask_origin, size_t mask_stride, size_t width, size_t height);
int rgpu_layer_blend_over_dev(rgpu_ctx* ctx, float* dst_dev, size_t dst_origin, size_t dst_stride, const float* src_dev,
                              size_t src_origin, size_t src_stride, size_t width, size_t height, int use_opacity, float opacity);
/* LinColor device image -> RGBA8 HOST image in one call (conversion kernel + a 4 B/pixel download): the export step that
 * follows `Scene::render` in every CLI run (`ImageOwned<LinColor>` -> RGBA8, src/image.rs:136-231, src/color.rs:164-175). */
int rgpu_download_rgba8(rgpu_ctx* ctx, const float* lin_dev, size_t n_pixels, uint8_t* rgba_host);

/* ---- batches of independent paths (BASELINE config 4; SURVEY §8e "batch of independent paths") -------------- */
/* Many paths in one flat encoding: `all` holds the subpaths of every path back to back and path i owns subpaths
 * [path_subpath_offsets[i], path_subpath_offsets[i + 1]) (n_paths + 1 offsets, the first 0, the last all->n_subpaths).
 * This is what a Rust caller builds from a `&[Path]` (src/path.rs:227-233) with one pass over `Path::subpaths()`. */

/* Upload n_paths paths with two copies; element i of the batch (rgpu_path_batch_get) is an ordinary device path that
 * rgpu_job::path may point to for as long as the batch lives. */
typedef struct rgpu_dpath_batch rgpu_dpath_batch;
int rgpu_path_upload_batch(rgpu_ctx* ctx, const rgpu_path* all, const uint32_t* path_subpath_offsets, size_t n_paths,
                           rgpu_dpath_batch** out);
const rgpu_dpath* rgpu_path_batch_get(const rgpu_dpath_batch* batch, size_t i);
void rgpu_path_batch_free(rgpu_ctx* ctx, rgpu_dpath_batch* batch);

/* Batch SVG path parse + `Path::bbox` + `fit_size` on the device: `text` holds n_paths SVG path strings back to
 * back (`d` attribute syntax), string i = bytes [text_offsets[i], text_offsets[i + 1]).  Every string is parsed by the
 * reference's grammar and number scanner (`SvgPathParser` src/svg.rs:241-421; scalars src/svg.rs:165-235: value =
 * (i64 mantissa as f64) * powi(10, exponent), not strtod) into `PathBuilder` semantics (src/path.rs:832-972: `line_to` drops
 * a line shorter than EPSILON, `close` returns to the subpath's first point, arcs become cubics), one thread per string, and
 * the result is an ordinary device path batch (rgpu_path_batch_get, rgpu_job::path).  Per path, `info[i]` (HOST array of
 * n_paths entries, may be NULL) receives `Path::bbox(identity)` (src/path.rs:428-431), the segment counts, the parse status
 * (`SvgParserError`, src/svg.rs:604-640; a failed path is empty) and — when `opt->fit_align >= 0` — `fit_size(bbox,
 * Size {fit_width, fit_height}, align)` (src/geometry.rs:490-516: the canvas size and the transform that fits the path
 * into it).  With info == NULL a parse error in any path fails the call (RGPU_ERR_INVALID, rgpu_last_error names the path,
 * the error kind and the byte offset).  Control points are bit-identical to the reference's except inside arcs (`A` / `a`:
 * sin / cos / tan / acos, a few ulp). */
enum { RGPU_PARSE_OK = 0, RGPU_PARSE_INVALID_CMD = 1, RGPU_PARSE_INVALID_SCALAR = 2, RGPU_PARSE_INVALID_FLAG = 3 };
enum { RGPU_ALIGN_MIN = 0, RGPU_ALIGN_MID = 1, RGPU_ALIGN_MAX = 2 }; /* `Align`, src/geometry.rs:298-305 */
typedef struct {
    double bbox[4];   /* min x, min y, max x, max y; valid when has_bbox */
    double fit_tr[6]; /* identity unless a fit was requested and has_bbox */
    uint32_t fit_width, fit_height;
    uint32_t n_points, n_segments, n_subpaths;
    int32_t status;        /* RGPU_PARSE_* */
    uint32_t error_offset; /* byte offset inside the path's string */
    int32_t has_bbox;      /* 0 for an empty path (`Path::bbox` returns None) */
    uint32_t n_curves, reserved;
} rgpu_parse_info;
typedef struct {
    uint32_t fit_width, fit_height; /* 0 = derive from the other side and the bbox's aspect ratio, as `fit_size` does */
    int32_t fit_align;              /* RGPU_ALIGN_*, or < 0: no fit */
} rgpu_parse_options;
int rgpu_parse_svg_batch(rgpu_ctx* ctx, const char* text, const uint32_t* text_offsets, size_t n_paths, const rgpu_parse_options* opt,
                         rgpu_dpath_batch** out, rgpu_parse_info* info);
/* Totals of a device path batch and its download in the flat batch encoding of rgpu_path_upload_batch:
 * points[2 * n_points], kinds[n_segments], subpath_offsets[n_subpaths + 1], closed[n_subpaths],
 * path_subpath_offsets[n_paths + 1]. */
int rgpu_path_batch_info(const rgpu_dpath_batch* batch, size_t* n_paths, uint32_t* n_points, uint32_t* n_segments, uint32_t* n_subpaths);
int rgpu_path_batch_download(rgpu_ctx* ctx, const rgpu_dpath_batch* batch, double* points, uint8_t* kinds, uint32_t* subpath_offsets,
                             uint8_t* closed, uint32_t* path_subpath_offsets);

/* A job list marshalled once and kept on the device: the steady state of a render loop that re-submits the same
 * batch (rgpu_render_batch re-derives and compares its tables on every call, which for 10^5 glyph jobs costs more
 * host time than the kernel takes).  `jobs` (and the paints they point to) must stay valid until rgpu_batch_free.
 * rgpu_batch_render is asynchronous like rgpu_render_batch; rgpu_batch_status / rgpu_sync apply. */
typedef struct rgpu_batch rgpu_batch;
int rgpu_batch_create(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags, rgpu_batch** out);
int rgpu_batch_render(rgpu_ctx* ctx, rgpu_batch* batch);
void rgpu_batch_free(rgpu_ctx* ctx, rgpu_batch* batch);

/* Output formats of the host-buffer batch calls */
enum {
    RGPU_OUT_LINCOLOR = 0, /* f32 x 4 premultiplied linear, 16 B per pixel: what `Path::fill` leaves in an `ImageOwned<LinColor>` */
    RGPU_OUT_RGBA8 = 1,    /* the same image after `From<LinColor> for RGBA` (src/color.rs:164-175), 4 B per pixel */
    RGPU_OUT_COVERAGE = 2  /* f32 coverage, dense form of `Rasterizer::mask_iter` (no paint), 4 B per pixel */
};
/* `ImageOwned::new_default(Size { width, height })` + `Path::fill(rasterizer, tr, fill_rule, paint, &mut img)`
 * (src/path.rs:492-507, src/rasterize.rs:70-115) for n_paths independent paths with HOST buffers: path i is filled at
 * trs[6 i .. 6 i + 6) (NULL = identity for all) onto its own fresh image, and the images are returned back to back in
 * `out_host` (n_paths x height x width pixels of `out_format`).  The batch runs in chunks: while chunk k renders,
 * chunk k - 1 is on its way to the host on a second stream (cudaMemcpyAsync into `out_host`, which should be pinned)
 * and chunk k + 1 is prepared on a helper thread.  RGPU_OUT_LINCOLOR with a plain solid paint: a share of every chunk
 * crosses PCIe as f32 coverage and host threads write colour * coverage — the multiplication the kernel does last — into
 * `out_host` (the same bytes; the share adapts to the host).  `paint` is ignored for RGPU_OUT_COVERAGE. */
int rgpu_fill_batch_host(rgpu_ctx* ctx, const rgpu_path* all, const uint32_t* path_subpath_offsets, size_t n_paths, const double* trs,
                         int fill_rule, const rgpu_paint* paint, uint32_t width, uint32_t height, int out_format, void* out_host);

/* ---- several GPUs of one box (no collective: every device returns its shard over its own PCIe link) --- */
/* One context + worker thread + streams per device; `devices` lists CUDA device ordinals (NULL = 0 .. n_devices-1). */
typedef struct rgpu_multi rgpu_multi;
int rgpu_multi_create(const int* devices, int n_devices, double flatness, rgpu_multi** out);
void rgpu_multi_destroy(rgpu_multi* m);
int rgpu_multi_device_count(const rgpu_multi* m);
const char* rgpu_multi_last_error(const rgpu_multi* m); /* m may be NULL: error of the last failed rgpu_multi_create */
/* rgpu_fill_batch_host sharded by path: device d takes a contiguous range of paths, cut so that the ranges hold equal
 * numbers of segments, and writes its images into its own region of `out_host` (disjoint; same result as one device). */
int rgpu_multi_fill_batch_host(rgpu_multi* m, const rgpu_path* all, const uint32_t* path_subpath_offsets, size_t n_paths,
                               const double* trs, int fill_rule, const rgpu_paint* paint, uint32_t width, uint32_t height,
                               int out_format, void* out_host);
/* `Rasterizer::mask` (src/rasterize.rs:299-311) of ONE path on a huge canvas, sharded by scanline bands: rows are
 * independent in the signed-difference rasterizer (src/rasterize.rs:421-469, 478-503), so the canvas is cut into n_bands
 * bands of rows (0 = one per device; cut points are multiples of 8 rows), device d takes the contiguous block of bands
 * [d * n_bands / n_devices, (d + 1) * n_bands / n_devices) as ONE job — equal rows are equal output bytes, which is what
 * bounds the raster kernel — flattens the path with the canvas transform, shifts the finished lines by the block's integer
 * row origin (the reference's own y clipping then crops exactly) and brings its rows into the host image (large blocks
 * run-coded: class byte per 64-pixel segment + literal edge segments over PCIe, rows rebuilt by host threads).  `img` is a dense width x height image of f32 (elem_size 4, the
 * device-native format) or f64 (elem_size 8, the trait's `Scalar`); it does not have to be zero on entry (every pixel is
 * written).  The result is bit-identical to the single-device rgpu_mask_f32 / rgpu_mask at every canvas size. */
int rgpu_multi_mask_banded_host(rgpu_multi* m, const rgpu_path* path, const double tr[6], int fill_rule, void* img, size_t elem_size,
                                size_t width, size_t height, uint32_t n_bands);
/* The same band decomposition on ONE context; also what every worker of rgpu_multi_mask_banded_host runs for its block.
 * Renders bands [band_first, band_first + band_count) of n_bands (consecutive bands are one job, flattened once) into their
 * rows of `img`; the other rows are not touched. */
int rgpu_mask_banded_host(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, void* img, size_t elem_size,
                          size_t width, size_t height, uint32_t n_bands, uint32_t band_first, uint32_t band_count);

/* ---- plumbing ----------------------------------------------------------------------------------- */
void* rgpu_stream(rgpu_ctx* ctx);  /* cudaStream_t the context launches on */
int rgpu_sync(rgpu_ctx* ctx);
int rgpu_device_alloc(rgpu_ctx* ctx, size_t bytes, void** out);
int rgpu_device_free(rgpu_ctx* ctx, void* p);
int rgpu_device_zero(rgpu_ctx* ctx, void* p, size_t bytes);
int rgpu_memcpy_h2d(rgpu_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int rgpu_memcpy_d2h(rgpu_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
/* pinned host memory for images handed to the host-buffer entry points (they also accept pageable memory) */
int rgpu_host_alloc(rgpu_ctx* ctx, size_t bytes, void** out);
int rgpu_host_free(rgpu_ctx* ctx, void* p);

#ifdef __cplusplus
}
#endif
#endif /* RASTERIZE_B200_H */
