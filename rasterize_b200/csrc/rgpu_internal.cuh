// rasterize_b200 — internal device-side structures shared by the kernels and the C-ABI glue.
//
// HBM layout:
//   points      double2[n_points]            path control points, as uploaded (f64, 16 B each)
//   items       uint2[n_segments+n_subpaths] one per segment plus one closing item per subpath
//   slot_counts u32[8*items+1]               lines produced by each (item, depth-3 subtree) slot
//   slot_offs   u32[8*items+1]               exclusive scan of slot_counts; last = total lines
//   lines       double4[n_lines]             flattened lines in the reference's order (x0,y0,x1,y1)
//   tile_counts u32[tiles+1] / tile_offs     per (job, band, chunk) tile reference counts / exclusive scan
//   bin_lines   double4[n_refs]              the lines of every tile, grouped by tile (order inside a tile is
//                                            irrelevant: accumulation is fixed-point, hence associative)
//   tile_state  u64[tiles*8]                 per-row tile totals / inclusive prefixes for the carry look-back
//   canvases    caller-owned                 f32 coverage or f32x4 LinColor
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rgpu {

constexpr int kSlotsPerItem = 8;        // depth-3 cut of every curve's subdivision tree
constexpr int kSlotDepth = 3;
constexpr int kMaxStack = 20;           // DFS stack depth below the slot root
constexpr uint32_t kItemClosing = 0x80000000u;
constexpr uint32_t kItemExplicitClosed = 0x40000000u;
constexpr uint32_t kItemIndexMask = 0x3fffffffu;

constexpr int kFixShift = 24;           // Q7.24 fixed-point winding cells
constexpr int kFixOne = 1 << kFixShift;
constexpr double kFixScale = 16777216.0;

constexpr int kMaxStops = 32;
constexpr int kStateRows = 8;             // rows per tile in the carry look-back state (chunked tiles have 8 rows)

enum JobMode : int { kModeMask = 0, kModeCoverage = 1, kModeFill = 2 };

struct PaintDev {
    int kind, linear_colors, spread, n_stops;
    double pixel_tr[6];          // (tr * [bbox.unit_transform *] paint.tr)^-1
    double p0x, p0y, p1x, p1y;   // linear: start/end ; radial: center/fcenter
    double dirx, diry;           // linear: (end-start)/|end-start|^2
    double r0, r1;               // radial: radius / fradius
    float solid[4];
    double stop_pos[kMaxStops];
    float stop_col[kMaxStops][4];
};

struct JobDev {
    double tr[6];
    const double2* pts;
    const uint2* items;
    uint32_t item_begin;   // first global item of this job
    uint32_t n_items;
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
};

struct Status {
    uint32_t nan_flag;
    uint32_t depth_flag;
    uint32_t lines_overflow;
    uint32_t refs_overflow;
    uint32_t n_lines;
    uint32_t n_refs;
    uint32_t pad[2];
};

// ---- launch wrappers (defined in the .cu files) -------------------------------------------------
void launch_flatten_count(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, uint32_t* slot_counts,
                          Status* status, cudaStream_t s);
void launch_flatten_emit(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, const uint32_t* slot_offs,
                         double4* lines, uint32_t lines_cap, Status* status, cudaStream_t s);
size_t scan_temp_bytes(uint32_t n);
void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, void* temp, size_t temp_bytes, cudaStream_t s);
// Unordered single-kernel flatten for the raster path: count + CTA-level reservation + emit (see flatten.cu).
// Writes status->n_lines (must be zero on entry) and, when line_job != nullptr, the job of every line.
void launch_flatten_fused(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, double4* lines, uint32_t* line_job,
                          uint32_t lines_cap, Status* status, cudaStream_t s);
// slot_offs == nullptr: lines came from launch_flatten_fused (count in status->n_lines, jobs in line_job)
void launch_bin_count(const JobDev* jobs, uint32_t n_jobs, const uint32_t* slot_offs, uint32_t total_slots, const uint32_t* line_job,
                      const double4* lines, uint32_t* tile_counts, int band_rows, int chunk_cols, Status* status, cudaStream_t s);
void launch_bin_fill(const JobDev* jobs, uint32_t n_jobs, const uint32_t* slot_offs, uint32_t total_slots, const uint32_t* line_job,
                     const double4* lines, const uint32_t* tile_offs, uint32_t total_tiles, uint32_t* tile_cursor, double4* bin_lines,
                     uint32_t refs_cap, int band_rows, int chunk_cols, Status* status, cudaStream_t s);
// tile geometry of the raster kernel variants
struct TileShape { int cw, th; };
TileShape raster_tile_shape(int variant);
// `ticket` is a zeroed device counter private to this launch (dynamic tile ids for the carry look-back);
// `tile_state` holds kMaxBandRows u64 words per tile, validated by `epoch` (no clearing between batches).
// `h_jobs` is the host copy of the job table: single-job launches pass their descriptor by value.
void launch_raster(int variant, const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first,
                   uint32_t n_tiles, const PaintDev* paints, const uint32_t* tile_offs, const double4* bin_lines,
                   unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, const Status* status, cudaStream_t s);
// Fused one-CTA-per-job pipeline for canvases of at most 64 x 64 visible pixels (small.cu)
bool small_canvas_eligible(uint32_t width, uint32_t height, int mode);
void launch_small_canvas(const JobDev* jobs, uint32_t job_first, uint32_t n_jobs, const PaintDev* paints, double thr, Status* status,
                         cudaStream_t s);
void launch_to_rgba8(const float4* lin, uchar4* out, size_t n, cudaStream_t s);
void launch_fill_color(float4* lin, size_t n, float4 color, cudaStream_t s);
void launch_f32_to_f64(const float* in, double* out, size_t n, cudaStream_t s);

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
