// K2 — binning of flattened lines into tiles: one bin per (job, scanline band, column chunk).
//
// Rows are independent in the signed-difference rasterizer (reference src/rasterize.rs:421-469: accumulation is
// per (line,row); :478-503: the scan runs along x within a row), so a line is referenced from every band of
// `band_rows` rows its y-range touches, and inside a band from every chunk of `chunk_cols` columns its cells can
// land in.  The row range is the reference's own:
//   first = floor(max(min_y, 0)),  end = min(H, ceil(max(max_y, 0)))     (src/rasterize.rs:414, 421)
// The column range is conservative (x of the line over the band's rows, clamped like the reference clamps to
// [0, width], one pixel of slack on both sides); the raster kernel drops what does not land in its tile.
// Lines with |dy| < EPSILON add nothing (src/rasterize.rs:400-403) and are dropped here.
#include "rgpu_internal.cuh"

namespace rgpu {

namespace {

constexpr double kEps = 2.220446049250313e-16;

// Lines come out of the flatten stage in path order, so neighbouring lanes almost always hit the same band:
// un-aggregated atomics serialise on one or two counters at a time (r1a profile: 95 warps stalled per issue).
// Each round, lanes holding the same band key elect a leader with __match_any_sync; the leader issues ONE
// atomic for the group and, for the fill pass, hands out consecutive slots by rank.
__device__ __forceinline__ uint32_t line_job_of(const JobDev* __restrict__ jobs, uint32_t n_jobs, const uint32_t* __restrict__ slot_offs,
                                                const uint32_t* __restrict__ line_job, uint32_t i) {
    if (n_jobs == 1) return 0;
    if (line_job) return line_job[i];
    return find_job(n_jobs, i, [&](uint32_t k) { return slot_offs[jobs[k].item_begin * kSlotsPerItem]; });
}

template <bool FILL>
__global__ void __launch_bounds__(256)
bin_kernel(const JobDev* __restrict__ jobs, uint32_t n_jobs, const uint32_t* __restrict__ slot_offs, uint32_t total_slots,
           const uint32_t* __restrict__ line_job, const double4* __restrict__ lines, uint32_t* __restrict__ tile_counts,
           const uint32_t* __restrict__ tile_offs, uint32_t total_tiles, double4* __restrict__ bin_lines, uint32_t refs_cap, int band_shift,
           int chunk_shift, Status* __restrict__ status) {
    if (status->lines_overflow | status->nan_flag | status->depth_flag) return;
    const uint32_t n_lines = slot_offs ? slot_offs[total_slots] : status->n_lines;
    if (!FILL) {
        if (slot_offs && blockIdx.x == 0 && threadIdx.x == 0) status->n_lines = n_lines;
    } else {
        const uint32_t n_refs = tile_offs[total_tiles];
        if (blockIdx.x == 0 && threadIdx.x == 0) status->n_refs = n_refs;
        if (n_refs > refs_cap) {
            if (blockIdx.x == 0 && threadIdx.x == 0) status->refs_overflow = 1u;
            return;
        }
    }
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < n_lines; i0 += stride) {  // warp-uniform bound
        const uint32_t i = i0 + lane;
        int b0 = 0, b1 = -1, n_chunks = 1;
        uint32_t tile_base = 0;
        double x0 = 0, y0 = 0, x1 = 0, y1 = 0, wc = 0;
        double4 l = make_double4(0, 0, 0, 0);
        if (i < n_lines) {
            l = lines[i];
            const uint32_t j = line_job_of(jobs, n_jobs, slot_offs, line_job, i);
            x0 = l.x; y0 = l.y; x1 = l.z; y1 = l.w;
            if (fabs(y0 - y1) >= kEps) {  // horizontal (or NaN) lines add no signed coverage
                const double H = (double)jobs[j].height;
                const double lo = fmin(y0, y1), hi = fmax(y0, y1);
                if (hi > 0.0 && lo < H) {
                    const double first = floor(fmax(lo, 0.0));
                    const double end = fmin(H, ceil(hi));
                    if (first < end) {
                        b0 = (int)first >> band_shift;
                        b1 = ((int)end - 1) >> band_shift;
                        tile_base = jobs[j].tile_begin;
                        n_chunks = (int)jobs[j].n_chunks;
                        wc = jobs[j].clamp_w;
                    }
                }
            }
        }
        // column range per band in f32 with 1.5 px of slack each side (conservative; the raster kernel is exact)
        const float fx0 = (float)x0, fy0 = (float)y0;
        const float fdxdy = (float)((x1 - x0) / (y1 - y0));
        const float fylo = (float)fmin(y0, y1), fyhi = (float)fmax(y0, y1), fwc = (float)wc;
        for (int b = b0;; b++) {
            const bool bvalid = b <= b1;
            if (__ballot_sync(0xffffffffu, bvalid) == 0) break;
            int c0 = 0, c1 = -1;
            if (bvalid) {
                if (n_chunks > 1) {
                    // x of the line at the top and bottom of its part inside this band, clamped like the reference
                    const float ya = fmaxf((float)(b << band_shift), fylo);
                    const float yb = fminf((float)((b + 1) << band_shift), fyhi);
                    const float xa = fx0 + (ya - fy0) * fdxdy, xb = fx0 + (yb - fy0) * fdxdy;
                    const float lo = fminf(fmaxf(fminf(xa, xb), 0.0f), fwc), hi = fminf(fmaxf(fmaxf(xa, xb), 0.0f), fwc);
                    c0 = max(0, (int)(lo - 1.5f) >> chunk_shift);
                    c1 = min(n_chunks - 1, (int)(hi + 2.5f) >> chunk_shift);
                    if (!(lo == lo) || !(hi == hi)) { c0 = 0; c1 = n_chunks - 1; }  // NaN from degenerate input: be conservative
                } else {
                    c0 = c1 = 0;
                }
            }
            for (int c = c0;; c++) {
                const bool valid = bvalid && c <= c1;
                const unsigned m = __ballot_sync(0xffffffffu, valid);
                if (m == 0) break;
                if (valid) {
                    const uint32_t key = tile_base + (uint32_t)b * (uint32_t)n_chunks + (uint32_t)c;
                    const unsigned peers = __match_any_sync(m, key);
                    const int leader = __ffs(peers) - 1;
                    uint32_t slot0 = 0;
                    if ((int)lane == leader) slot0 = atomicAdd(&tile_counts[key], (uint32_t)__popc(peers));
                    if (FILL) {
                        slot0 = __shfl_sync(peers, slot0, leader);
                        const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
                        bin_lines[tile_offs[key] + slot0 + rank] = l;  // the line itself: the raster kernel reads its bin coalesced
                    }
                }
            }
        }
    }
}

}  // namespace

static inline uint32_t line_grid(cudaStream_t) { return 148 * 8; }
static inline int log2i(int v) {
    int s = 0;
    while ((1 << s) < v) s++;
    return s;  // tile shapes are powers of two
}

void launch_bin_count(const JobDev* jobs, uint32_t n_jobs, const uint32_t* slot_offs, uint32_t total_slots, const uint32_t* line_job,
                      const double4* lines, uint32_t* tile_counts, int band_rows, int chunk_cols, Status* status, cudaStream_t s) {
    bin_kernel<false><<<line_grid(s), 256, 0, s>>>(jobs, n_jobs, slot_offs, total_slots, line_job, lines, tile_counts, nullptr, 0, nullptr,
                                                   0, log2i(band_rows), log2i(chunk_cols), status);
}

void launch_bin_fill(const JobDev* jobs, uint32_t n_jobs, const uint32_t* slot_offs, uint32_t total_slots, const uint32_t* line_job,
                     const double4* lines, const uint32_t* tile_offs, uint32_t total_tiles, uint32_t* tile_cursor, double4* bin_lines,
                     uint32_t refs_cap, int band_rows, int chunk_cols, Status* status, cudaStream_t s) {
    bin_kernel<true><<<line_grid(s), 256, 0, s>>>(jobs, n_jobs, slot_offs, total_slots, line_job, lines, tile_cursor, tile_offs, total_tiles,
                                                  bin_lines, refs_cap, log2i(band_rows), log2i(chunk_cols), status);
}

}  // namespace rgpu
