// K2 — binning of flattened lines into (job, scanline band) bins (the scans live in scan.cu).
//
// Rows are independent in the signed-difference rasterizer (reference src/rasterize.rs:421-469: accumulation is
// per (line,row); :478-503: the scan runs along x within a row), so a line is referenced once from every band
// of `band_rows` rows its y-range touches.  The row range is the reference's own:
//   first = floor(max(min_y, 0)),  end = min(H, ceil(max(max_y, 0)))     (src/rasterize.rs:414, 421)
// Lines with |dy| < EPSILON add nothing (src/rasterize.rs:400-403) and are dropped here.
#include "rgpu_internal.cuh"

namespace rgpu {

namespace {

constexpr double kEps = 2.220446049250313e-16;

struct LineBands {
    uint32_t job;
    int b0, b1;  // inclusive band range, b1 < b0 when the line touches no row
};

__device__ __forceinline__ LineBands line_bands(const JobDev* __restrict__ jobs, uint32_t n_jobs, const uint32_t* __restrict__ slot_offs,
                                                uint32_t i, const double4 l, int band_rows) {
    LineBands r;
    r.job = find_job(n_jobs, i, [&](uint32_t k) { return slot_offs[jobs[k].item_begin * kSlotsPerItem]; });
    r.b0 = 0;
    r.b1 = -1;
    double y0 = l.y, y1 = l.w;
    if (!(fabs(y0 - y1) >= kEps)) return r;  // horizontal (or NaN) line: no signed coverage
    double H = (double)jobs[r.job].height;
    double lo = fmin(y0, y1), hi = fmax(y0, y1);
    if (!(hi > 0.0) || !(lo < H)) return r;
    double first = floor(fmax(lo, 0.0));
    double end = fmin(H, ceil(hi));  // hi > 0 here
    if (!(first < end)) return r;
    r.b0 = (int)first / band_rows;
    r.b1 = ((int)end - 1) / band_rows;
    return r;
}

__global__ void __launch_bounds__(256)
bin_count_kernel(const JobDev* __restrict__ jobs, uint32_t n_jobs, const uint32_t* __restrict__ slot_offs, uint32_t total_slots,
                 const double4* __restrict__ lines, uint32_t* __restrict__ band_counts, int band_rows, Status* __restrict__ status) {
    if (status->lines_overflow | status->nan_flag | status->depth_flag) return;
    uint32_t n_lines = slot_offs[total_slots];
    if (blockIdx.x == 0 && threadIdx.x == 0) status->n_lines = n_lines;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += gridDim.x * blockDim.x) {
        LineBands lb = line_bands(jobs, n_jobs, slot_offs, i, lines[i], band_rows);
        uint32_t base = jobs[lb.job].band_begin;
        for (int b = lb.b0; b <= lb.b1; b++) atomicAdd(&band_counts[base + b], 1u);
    }
}

__global__ void __launch_bounds__(256)
bin_fill_kernel(const JobDev* __restrict__ jobs, uint32_t n_jobs, const uint32_t* __restrict__ slot_offs, uint32_t total_slots,
                const double4* __restrict__ lines, const uint32_t* __restrict__ band_offs, uint32_t total_bands,
                uint32_t* __restrict__ band_cursor, uint32_t* __restrict__ refs, uint32_t refs_cap, int band_rows,
                Status* __restrict__ status) {
    if (status->lines_overflow | status->nan_flag | status->depth_flag) return;
    uint32_t n_refs = band_offs[total_bands];
    if (blockIdx.x == 0 && threadIdx.x == 0) status->n_refs = n_refs;
    if (n_refs > refs_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) status->refs_overflow = 1u;
        return;
    }
    uint32_t n_lines = slot_offs[total_slots];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += gridDim.x * blockDim.x) {
        LineBands lb = line_bands(jobs, n_jobs, slot_offs, i, lines[i], band_rows);
        uint32_t base = jobs[lb.job].band_begin;
        for (int b = lb.b0; b <= lb.b1; b++) {
            uint32_t slot = atomicAdd(&band_cursor[base + b], 1u);
            refs[band_offs[base + b] + slot] = i;
        }
    }
}

}  // namespace

static inline uint32_t line_grid(cudaStream_t) { return 148 * 8; }

void launch_bin_count(const JobDev* jobs, uint32_t n_jobs, const uint32_t* slot_offs, uint32_t total_slots, const double4* lines,
                      uint32_t* band_counts, int band_rows, Status* status, cudaStream_t s) {
    bin_count_kernel<<<line_grid(s), 256, 0, s>>>(jobs, n_jobs, slot_offs, total_slots, lines, band_counts, band_rows, status);
}

void launch_bin_fill(const JobDev* jobs, uint32_t n_jobs, const uint32_t* slot_offs, uint32_t total_slots, const double4* lines,
                     const uint32_t* band_offs, uint32_t total_bands, uint32_t* band_cursor, uint32_t* refs, uint32_t refs_cap,
                     int band_rows, Status* status, cudaStream_t s) {
    bin_fill_kernel<<<line_grid(s), 256, 0, s>>>(jobs, n_jobs, slot_offs, total_slots, lines, band_offs, total_bands, band_cursor,
                                                 refs, refs_cap, band_rows, status);
}

}  // namespace rgpu
