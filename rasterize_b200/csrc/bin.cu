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
           const uint32_t* __restrict__ line_job, const double4* __restrict__ lines, uint32_t* __restrict__ band_counts,
           const uint32_t* __restrict__ band_offs, uint32_t total_bands, uint32_t* __restrict__ refs, uint32_t refs_cap, int band_rows,
           Status* __restrict__ status) {
    if (status->lines_overflow | status->nan_flag | status->depth_flag) return;
    const uint32_t n_lines = slot_offs ? slot_offs[total_slots] : status->n_lines;
    if (!FILL) {
        if (slot_offs && blockIdx.x == 0 && threadIdx.x == 0) status->n_lines = n_lines;
    } else {
        const uint32_t n_refs = band_offs[total_bands];
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
        int b0 = 0, b1 = -1;
        uint32_t base = 0;
        if (i < n_lines) {
            const double4 l = lines[i];
            const uint32_t j = line_job_of(jobs, n_jobs, slot_offs, line_job, i);
            const double y0 = l.y, y1 = l.w;
            if (fabs(y0 - y1) >= kEps) {  // horizontal (or NaN) lines add no signed coverage
                const double H = (double)jobs[j].height;
                const double lo = fmin(y0, y1), hi = fmax(y0, y1);
                if (hi > 0.0 && lo < H) {
                    const double first = floor(fmax(lo, 0.0));
                    const double end = fmin(H, ceil(hi));
                    if (first < end) {
                        b0 = (int)first / band_rows;
                        b1 = ((int)end - 1) / band_rows;
                        base = jobs[j].band_begin;
                    }
                }
            }
        }
        for (int b = b0;; b++) {
            const bool valid = b <= b1;
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (m == 0) break;
            if (valid) {
                const uint32_t key = base + (uint32_t)b;
                const unsigned peers = __match_any_sync(m, key);
                const int leader = __ffs(peers) - 1;
                uint32_t slot0 = 0;
                if ((int)lane == leader) slot0 = atomicAdd(&band_counts[key], (uint32_t)__popc(peers));
                if (FILL) {
                    slot0 = __shfl_sync(peers, slot0, leader);
                    const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
                    refs[band_offs[key] + slot0 + rank] = i;
                }
            }
        }
    }
}

}  // namespace

static inline uint32_t line_grid(cudaStream_t) { return 148 * 8; }

void launch_bin_count(const JobDev* jobs, uint32_t n_jobs, const uint32_t* slot_offs, uint32_t total_slots, const uint32_t* line_job,
                      const double4* lines, uint32_t* band_counts, int band_rows, Status* status, cudaStream_t s) {
    bin_kernel<false><<<line_grid(s), 256, 0, s>>>(jobs, n_jobs, slot_offs, total_slots, line_job, lines, band_counts, nullptr, 0, nullptr,
                                                   0, band_rows, status);
}

void launch_bin_fill(const JobDev* jobs, uint32_t n_jobs, const uint32_t* slot_offs, uint32_t total_slots, const uint32_t* line_job,
                     const double4* lines, const uint32_t* band_offs, uint32_t total_bands, uint32_t* band_cursor, uint32_t* refs,
                     uint32_t refs_cap, int band_rows, Status* status, cudaStream_t s) {
    bin_kernel<true><<<line_grid(s), 256, 0, s>>>(jobs, n_jobs, slot_offs, total_slots, line_job, lines, band_cursor, band_offs, total_bands,
                                                  refs, refs_cap, band_rows, status);
}

}  // namespace rgpu
