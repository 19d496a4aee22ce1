// rasterize_b200 — internal device-side structures shared by the kernels and the C-ABI glue.
//
// HBM layout:
//   points      double2[n_points]            path control points, as uploaded (f64, 16 B each)
//   items       uint2[n_segments+n_subpaths] one per segment plus one closing item per subpath
//   slot_counts u32[8*items+1]               lines produced by each (item, depth-3 subtree) slot
//   slot_offs   u32[8*items+1]               exclusive scan of slot_counts; last = total lines
//   lines       double4[n_lines]             flattened lines in the reference's order (x0,y0,x1,y1)
//   tile_counts u32[tiles]                   lines each (job, band, chunk) tile wants (self-cleaning: zeroed by the raster CTA)
//   bin_lines   double4[tiles * bin_cap]     fixed-capacity bin per tile (order inside a tile is irrelevant: accumulation
//                                            is fixed-point, hence associative); two-pass fallback: packed, with tile_offs
//   tile_state  u64[tiles*16]                per-row tile totals / inclusive prefixes for the carry look-back
//   canvases    caller-owned                 f32 coverage or f32x4 LinColor
#pragma once
#include "stroke_units.hpp"
#include "parse_tables.hpp"
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rgpu {

constexpr int kSlotsPerItem = 8;        // depth-3 cut of every curve's subdivision tree
constexpr int kSlotDepth = 3;
constexpr int kMaxStack = 20;           // DFS stack depth below the slot root
constexpr uint32_t kItemClosing = 0x80000000u;
constexpr uint32_t kItemExplicitClosed = 0x40000000u;
constexpr uint32_t kItemIndexMask = 0x3fffffffu;

// Winding cells are 32-bit fixed point, Q7.24 by default: 6e-8 of a pixel, windings in [-128, 128).  The integer
// arithmetic wraps modulo 256 windings, which EvenOdd cannot see (256 is even) and NonZero sees only for a true winding
// within 1 of a non-zero multiple of 256.  The row scans watch for |winding| >= kWindingGuard under NonZero; a batch that
// trips the guard is re-run by the *_sync entry points in Q13.18 (4e-6 of a pixel, windings in [-8192, 8192)).
constexpr int kFixShift = 24;
constexpr int kFixShiftWide = 18;
constexpr int kWindingGuard = 120;

constexpr int kMaxStops = 32;
constexpr int kStateRows = 16;            // rows per tile in the carry look-back state (chunked tiles have 8 rows)

enum JobMode : int { kModeMask = 0, kModeCoverage = 1, kModeFill = 2, kModeRender = 3 };

struct PaintDev {
    int kind, linear_colors, spread, n_stops;
    double pixel_tr[6];          // (tr * [bbox.unit_transform *] paint.tr)^-1
    double p0x, p0y, p1x, p1y;   // linear: start/end ; radial: center/fcenter
    double dirx, diry;           // linear: (end-start)/|end-start|^2
    double r0, r1;               // radial: radius / fradius
    float solid[4];
    double stop_pos[kMaxStops];
    float stop_col[kMaxStops][4];
    // per-paint constants hoisted out of the per-pixel evaluation (filled by the host, context.cu: build_paint)
    double stop_inv[kMaxStops];  // 1 / (stop_pos[i] - stop_pos[i-1]) for i >= 1
    double lin_a, lin_b, lin_c;  // linear: t = (x + 0.5) * lin_a + (y + 0.5) * lin_b + lin_c (pixel_tr and dir folded together)
    double rad_cdx, rad_cdy, rad_rd, rad_a;  // radial: c - fc, r - fr, cd.cd - rd^2 (src/grad.rs:361-372)
    double rad_inv2a;                        // 1 / (2 a)
};

struct JobDev {
    double tr[6];
    const double2* pts;
    const uint2* items;
    uint32_t item_begin;   // first global item of this job
    uint32_t n_items;
    const uint2* items_packed;  // the same items, curves first (raster path: flatten_device.cuh slot_setup_packed)
    uint32_t n_curves;          // curve items of this job
    uint32_t thread_begin;      // first thread of this job in the packed grid
    uint32_t band_begin;   // first global band
    uint32_t n_bands;
    uint32_t n_chunks;
    uint32_t tile_begin;   // first global tile (band-major, chunk fastest)
    int32_t width_out;     // columns written
    int32_t height;
    double clamp_w;        // reference `width` of signed_difference_line: img.width - 1
    int32_t rule, mode, close, paint_index;
    void* canvas;
    unsigned long long origin, row_stride;  // elements
    // scene batches (scene.cu): the job's tile grid is aligned with the LAYER's tiles.  (ox, oy) = position of the job's
    // window inside its first layer tile, (sc0, sb0) = that tile's chunk / band index in the layer.  All zero otherwise.
    int32_t ox, oy, sc0, sb0;
    int32_t fix_shift;     // fixed-point fraction bits of this batch's winding cells (kFixShift or kFixShiftWide)
    int32_t pad_;
    // Row origin of a band job (rgpu_mask_banded_host): the path is flattened with the canvas transform — the very lines of
    // the unsharded job — and this (integer-valued) origin is subtracted from the finished lines' y before they are binned,
    // instead of folding a translate(0, -y0) into `tr`, which would round every transformed control point differently.  0 otherwise.
    double y_org;
};

struct Status {
    uint32_t nan_flag;
    uint32_t depth_flag;
    uint32_t lines_overflow;
    uint32_t refs_overflow;
    uint32_t n_lines;
    uint32_t n_refs;
    uint32_t bin_max;   // fixed-capacity bins: largest per-tile line count seen (> capacity => refs_overflow)
    uint32_t winding_flag;  // a NonZero winding reached kWindingGuard: the 32-bit cells may wrap (see kFixShift)
};

// tile geometry of the raster kernel variants
struct TileShape { int cw, th; };

// ---- launch wrappers (defined in the .cu files) -------------------------------------------------
void launch_flatten_count(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, uint32_t* slot_counts,
                          Status* status, cudaStream_t s);
void launch_flatten_emit(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, const uint32_t* slot_offs,
                         double4* lines, uint32_t lines_cap, Status* status, cudaStream_t s);
size_t scan_temp_bytes(uint32_t n);
// `temp` must be zero on entry; pass temp_is_zero = true when the caller has already cleared it
void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, void* temp, size_t temp_bytes, cudaStream_t s,
                           bool temp_is_zero = false);
// Flatten fused with binning, two-pass fallback (fixed bins over budget): pass 0 walks every slot and counts lines per
// tile (and in total, into status->n_lines); after an exclusive scan of the tile counts, pass 1 walks again and writes
// every line straight into the packed bins of the tiles it touches.  No global line buffer.
void launch_flatten_bin_count(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, uint32_t* tile_counts,
                              int band_rows, int chunk_cols, Status* status, cudaStream_t s);
void launch_flatten_bin_emit(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, const uint32_t* tile_offs,
                             uint32_t total_tiles, uint32_t* tile_cursor, double4* bin_lines, uint32_t refs_cap, int band_rows,
                             int chunk_cols, Status* status, cudaStream_t s);
// Single-pass variant: every tile owns a fixed bin of `bin_cap` lines at bin_lines[tile * bin_cap]; tile_counts ends
// up holding the number of lines each tile wanted.  A tile that wants more than bin_cap sets refs_overflow (and
// bin_max) and the host re-runs with a larger capacity or with the exact two-pass scheme.
// `h_jobs` is the host copy of the job table: a single-job batch passes its descriptor by value and `jobs` is not read.
// `next_status` (may be NULL) is cleared for the following batch.
// The kernel runs on the PACKED grid: `total_threads` = sum of the jobs' thread counts, job j's threads start at
// JobDev::thread_begin = multiples of 32, (n_curves << depth) curve-slot threads then one per line item;
// `depth` = flatten_cut_depth(total items of the batch).
int flatten_cut_depth(uint32_t total_items);
void launch_flatten_bin_fixed(const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t total_threads, int depth, double thr,
                              uint32_t* tile_counts, double4* bin_lines, uint32_t bin_cap, int band_rows, int chunk_cols, Status* status,
                              Status* next_status, cudaStream_t s);
TileShape raster_tile_shape(int variant);
// ---- mask_iter (compact.cu): the pixels with coverage != 0 of a dense n_pixels canvas as rgpu_pixel records, row-major ----
uint32_t pixel_blocks(size_t n_pixels);
void launch_pixel_count(const float* cov, size_t n_pixels, uint32_t* counts, cudaStream_t s);   // counts[pixel_blocks(n_pixels)]
void launch_pixel_emit(const float* cov, size_t n_pixels, size_t width, const uint32_t* offs, void* out, size_t cap, cudaStream_t s);
// run-coded rows (compact.cu): class byte per 64-pixel segment of a dense f32 image (0: all +0.0f, 1: all 1.0f, 2: literal),
// literal count per row, and — after an exclusive scan of the counts — the literals packed in (row, segment) order
uint32_t runcode_segments(size_t width);
void launch_seg_classify(const float* img, size_t width, uint32_t rows, unsigned char* cls, uint32_t* row_cnt, cudaStream_t s);
void launch_seg_emit(const float* img, size_t width, uint32_t rows, const unsigned char* cls, const uint32_t* row_off, float* lits, cudaStream_t s);
// ---- stroke (stroke.cu; the unit table is described in stroke_units.hpp) ----
struct StrokeStyleDev {
    double width, miter_limit;
    int join, cap;
};
// cnt / off: four arrays of stroke_count_stride(n_units) words each (segments, points, curves, closed contours)
void launch_stroke_pieces(const StrokeUnit* units, uint32_t n_units, const double2* pts, const StrokeStyleDev& st, uint32_t* cnt,
                          void* first, void* last, cudaStream_t s);
void launch_stroke_units(bool emit, const StrokeUnit* units, uint32_t n_units, const double2* pts, const StrokeStyleDev& st, uint32_t* cnt,
                         const void* first, const void* last, const uint32_t* off, double2* out_pts, uint2* out_items, uint2* out_packed,
                         cudaStream_t s);
size_t stroke_piece_bytes();
// ---- batch SVG parse (parse.cu; tables and the host-side plan are in parse_plan.hpp) ----
void launch_parse_count(const uint8_t* text, const uint32_t* chunk_off, uint32_t n_chunks, const ParseFit& fit, ParseInfoDev* info,
                        cudaStream_t s);
void launch_parse_emit(const uint8_t* text, const uint32_t* chunk_off, uint32_t n_chunks, const ParseEmitBase* bases, double2* out_pts,
                       uint2* out_items, uint2* out_packed, cudaStream_t s);
void parse_plan_chunks_host(const char* text, const uint32_t* text_off, uint32_t n_paths, std::vector<uint32_t>& chunk_off,
                            std::vector<uint32_t>& chunk_first);
void parse_merge_chunks_host(const ParseInfoDev* info, const std::vector<uint32_t>& chunk_off, const std::vector<uint32_t>& chunk_first,
                             const uint32_t* text_off, uint32_t n_paths, const ParseFit& fit, ParseInfoDev* path_info,
                             std::vector<ParseEmitBase>& bases, std::vector<uint32_t>& item_off, uint32_t& total_pts);

// `ticket` is a zeroed device counter private to this launch (dynamic tile ids for the carry look-back);
// `tile_state` holds kMaxBandRows u64 words per tile, validated by `epoch` (no clearing between batches).
// `h_jobs` is the host copy of the job table: single-job launches pass their descriptor by value.
// `pdl` (programmatic dependent launch): 1 = the launch directly follows the flatten kernel in the stream and may overlap
// its tail; the kernel waits for it before reading anything it wrote.  2 = the launch follows the previous fill of the same
// canvas in an ordered batch: it accumulates its tiles while that fill composites and waits only before compositing.  0 = plain.
// `zero_early`: clear every tile's cells before its id is known (pays off when most tiles hold lines).
// bin_cap == 0: tile t's lines are bin_lines[tile_offs[t] .. tile_offs[t+1]); bin_cap > 0: fixed bins, tile t's lines
// are bin_lines[t*bin_cap .. t*bin_cap + tile_offs[t]) (tile_offs then holds the per-tile COUNTS).
// With fixed bins every tile clears its own counter after reading it and the last ticket resets the ticket counter, so
// the next batch finds them zero without a memset.
void launch_raster(int variant, const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first,
                   uint32_t n_tiles, const PaintDev* paints, uint32_t* tile_offs, uint32_t bin_cap, const double4* bin_lines,
                   unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, Status* status, bool zero_early, int pdl, cudaStream_t s);
// Scene compositor (scene.cu): every FILL job of a layer in one launch, one CTA per 256 x 8 LAYER tile, fills blended in
// submission order inside the CTA.  Jobs carry layer-aligned tile grids (JobDev::ox/oy/sc0/sb0); fixed bins only.
struct SceneArgs {
    float4* layer;              // dense W x H LinColor layer (row pitch = width)
    uchar4* rgba;               // optional RGBA8 export of the finished layer (same geometry), or NULL
    uint32_t width, height;
    uint32_t n_bands, n_chunks; // layer tile grid
    float bg[4];                // fresh != 0: every pixel starts from this colour (`Layer::new`), else from the layer's content
    int fresh, store_lin;
    // jobs whose window reaches layer band B, in submission order: band_jobs[band_offs[B] .. band_offs[B + 1])
    // (a CTA walks only those instead of the whole job table; filled in by the host, context.cu)
    const uint32_t* band_offs;
    const uint32_t* band_jobs;
    // bands in the order the tickets take them inside a column of tiles: heaviest first (bands are independent of one
    // another; only the chunks of a band are chained), so the last CTAs of the launch are the cheap ones
    const uint32_t* band_order;
};
TileShape scene_tile_shape();
void launch_scene(const JobDev* jobs, uint32_t n_jobs, const PaintDev* paints, uint32_t* tile_offs, uint32_t bin_cap, const double4* bin_lines,
                  unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, Status* status, const SceneArgs& sc, bool pdl,
                  cudaStream_t s);
// Fused one-CTA-per-job pipeline for canvases of at most 64 x 64 visible pixels (small.cu)
bool small_canvas_eligible(uint32_t width, uint32_t height, int mode);
// `gradients`: some job of the launch has a gradient paint (selects the kernel variant that carries the paint evaluation)
void launch_small_canvas(const JobDev* jobs, uint32_t job_first, uint32_t n_jobs, const PaintDev* paints, double thr, Status* status,
                         bool gradients, cudaStream_t s);
void launch_to_rgba8(const float4* lin, uchar4* out, size_t n, cudaStream_t s);
void launch_fill_color(float4* lin, size_t n, float4 color, cudaStream_t s);
void launch_f32_to_f64(const float* in, double* out, size_t n, cudaStream_t s);
// Layer::compose on device (compose.cu): strides in pixels, pointers already offset to the intersection rectangle
void launch_scale_by_mask(float4* lin, unsigned long long lin_stride, const float* mask, unsigned long long mask_stride, uint32_t width,
                          uint32_t height, cudaStream_t s);
void launch_blend_over(float4* dst, unsigned long long dst_stride, const float4* src, unsigned long long src_stride, uint32_t width,
                       uint32_t height, bool use_opacity, float opacity, cudaStream_t s);

// Tiles a flattened line may touch: every band of 2^band_shift rows its y-range covers (the reference's own row
// range, src/rasterize.rs:414, 421) and, inside a band, every chunk of 2^chunk_shift columns its cells can land in
// (x over the band's rows, clamped like the reference clamps to [0, width], 1.5 px of slack; the raster kernel is
// exact and drops what is not in its tile).  Lines with |dy| < EPSILON add nothing (src/rasterize.rs:400-403).
//
// band_range: first band (layer-aligned index) and number of bands of a line, 0 = the line adds nothing.
__device__ __forceinline__ int band_range(const JobDev& job, double y0, double y1, int band_shift, int& b0) {
    b0 = 0;
    if (!(fabs(y0 - y1) >= 2.220446049250313e-16)) return 0;
    const double H = (double)job.height;
    const double lo = fmin(y0, y1), hi = fmax(y0, y1);
    if (!(hi > 0.0) || !(lo < H)) return 0;
    const double first = floor(fmax(lo, 0.0));
    const double end = fmin(H, ceil(hi));
    if (!(first < end)) return 0;
    b0 = ((int)first + job.oy) >> band_shift;
    return ((((int)end - 1 + job.oy) >> band_shift) - b0) + 1;
}

// The tiles of ONE band b (as counted by band_range) that the line may touch: calls f(global tile index) for each.
// The x range over the band's rows is evaluated in f64 — at y ~ 32768 an f32 row coordinate is 2e-3 off, which a shallow
// line's slope would turn into many pixels — with the slope itself from an f32 division (relative error 1e-7 of at
// most the line's x extent: far inside the 1.5 px of slack).
template <class F>
__device__ __forceinline__ void for_band_tiles(const JobDev& job, double x0, double y0, double x1, double y1, int b, int band_shift,
                                               int chunk_shift, F f) {
    const int n_chunks = (int)job.n_chunks;
    if (n_chunks == 1) {
        f(job.tile_begin + (uint32_t)b);
        return;
    }
    const int oy = job.oy, ox = job.ox;
    const double lo = fmin(y0, y1), hi = fmax(y0, y1);
    const double dxdy = (double)__fdividef((float)(x1 - x0), (float)(y1 - y0));
    const double ya = fmax((double)((b << band_shift) - oy), lo);
    const double yb = fmin((double)(((b + 1) << band_shift) - oy), hi);
    const double xa = fma(ya - y0, dxdy, x0), xb = fma(yb - y0, dxdy, x0);
    const double xl = fmin(fmax(fmin(xa, xb), 0.0), job.clamp_w), xh = fmin(fmax(fmax(xa, xb), 0.0), job.clamp_w);
    int c0 = max(0, (__double2int_rd(xl - 1.5) + ox) >> chunk_shift);
    int c1 = min(n_chunks - 1, (__double2int_rd(xh + 2.5) + ox) >> chunk_shift);
    if (!(xl == xl) || !(xh == xh)) { c0 = 0; c1 = n_chunks - 1; }  // NaN from degenerate input: be conservative
    for (int c = c0; c <= c1; c++) f(job.tile_begin + (uint32_t)b * (uint32_t)n_chunks + (uint32_t)c);
}

template <class F>
__device__ __forceinline__ void for_each_tile(const JobDev& job, double x0, double y0, double x1, double y1, int band_shift,
                                              int chunk_shift, F f) {
    int b0;
    const int span = band_range(job, y0, y1, band_shift, b0);
    for (int b = b0; b < b0 + span; b++) for_band_tiles(job, x0, y0, x1, y1, b, band_shift, chunk_shift, f);
}

// upper_bound-style search: largest j with begin[j] <= v, over a strided member of JobDev
template <class F>
__device__ __forceinline__ uint32_t find_job(uint32_t n_jobs, uint32_t v, F begin_of) {
    uint32_t lo = 0, hi = n_jobs;  // invariant: begin_of(lo) <= v < begin_of(hi) (hi == n_jobs is +inf)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (begin_of(mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

}  // namespace rgpu
