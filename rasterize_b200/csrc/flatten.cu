// K1 — curve flattening on device.
//
// Replaces `Path::flatten` / `PathFlattenIter` (reference src/path.rs:418-425, 744-795) and the curve maths it
// calls: `Transform::apply` (src/geometry.rs:363-367), `flatness` (src/curve.rs:233-235, 413-417, 692-697),
// `split` (src/curve.rs:43-45, 449-456, 731-747), `has_nans` (src/curve.rs:1081-1089).
//
// Line counts must be IDENTICAL to the reference, which depends on bit-exact f64 evaluation of the strict
// `flatness() < 16*tol^2` predicate.  Every arithmetic step below therefore uses the round-to-nearest
// intrinsics (__dmul_rn/__dadd_rn/__dsub_rn): they are never contracted into FMAs, whatever -fmad says, and the
// operand order follows the reference expression trees exactly.
//
// Parallelisation: the subdivision tree of every item is cut at depth 3 (8 "slots"; depth 5 = 32 slots for small
// batches); each slot's subtree is walked depth-first by its own thread with an explicit stack of pending right halves.
//   * `rgpu_flatten` (ordered): a count pass sizes every slot, an exclusive scan places them, an emit pass writes
//     the lines in the reference's order.
//   * raster path: the walk is fused with binning (flatten_bin_kernel), no global line buffer.  Default (PASS 2): ONE
//     walk; leaves are parked in a per-warp shared queue and then appended to the fixed-capacity bin of every tile they
//     touch.  Fallback (PASS 0 / 1): count lines per tile, scan, walk again and write packed bins.
#include "flatten_device.cuh"

#include <cstdlib>

#ifndef RGPU_FLAT_MINB
#define RGPU_FLAT_MINB 7  // 72 registers, no spills; with the packed grid the C2 launch is ~630 CTAs (one wave at any of 5..10 CTAs per SM)
#endif

namespace rgpu {

namespace {

using namespace fl;

constexpr int kWarpQueue = 96;                 // leaves a warp parks in shared memory before binning them (flatten_bin_kernel<2>)
constexpr int kDeepCutDepth = 5;              // 32 slots per item for small batches (see launch_flatten_bin_fixed)
constexpr uint32_t kDeepCutMaxItems = 8192;
constexpr int kDeepestCutDepth = 7;           // 128 slots per item for tiny batches (one outline, one scene)
constexpr uint32_t kDeepestCutMaxItems = 512;

// Ordered two-pass form (count, [scan], emit): lines land in the reference's exact order — `Path::flatten` parity.
template <bool EMIT>
__global__ void __launch_bounds__(128)
flatten_kernel(const JobDev* __restrict__ jobs, uint32_t n_jobs, uint32_t total_items, double thr,
               uint32_t* __restrict__ slot_counts, const uint32_t* __restrict__ slot_offs, double4* __restrict__ lines,
               uint32_t lines_cap, Status* __restrict__ status) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total_slots = total_items * kSlotsPerItem;
    if (t >= total_slots) {
        if (!EMIT && t == total_slots) slot_counts[t] = 0;  // sentinel so the scan yields the total
        return;
    }
    uint32_t out = 0, out_end = 0;
    if (EMIT) {
        out = slot_offs[t];
        out_end = slot_offs[t + 1];
        if (out_end == out) return;
        if (out_end > lines_cap) {  // capacity exceeded: flag and skip (host grows the buffer and re-runs)
            atomicExch(&status->lines_overflow, 1u);
            return;
        }
    }
    SlotCtx c;
    if (!slot_setup(jobs, n_jobs, t, thr, c, status)) {
        if (!EMIT) slot_counts[t] = 0;
        return;
    }
    if (EMIT) {
        slot_walk(c, thr, status, [&](double x0, double y0, double x1, double y1) {
            if (out < out_end) lines[out] = make_double4(x0, y0, x1, y1);
            out++;
        });
    } else {
        slot_counts[t] = slot_walk(c, thr, status, [](double, double, double, double) {});
    }
}

// Flatten fused with binning.  PASS 0: count lines per tile; PASS 1: write lines into their bins.
// Tried and measured no faster on C2 / C5 (round 1): CTA-level shared-memory aggregation of the tile counters; deferring
// the bin store behind the next leaf; and a warp work-sharing schedule (32 lanes on a shared LIFO of pending nodes in
// shared memory, critical path = subdivision depth) — its tail is set by runs of heavy curves landing in one warp, and
// with batches small enough to avoid that (strided, 8 items) it ties this kernel (33 us on C2); only C5 gained (31 -> 22 us).
template <int PASS, int DEPTH>
__global__ void __launch_bounds__(128, RGPU_FLAT_MINB)
flatten_bin_kernel(const JobDev* __restrict__ jobs_in, uint32_t n_jobs, const __grid_constant__ JobDev one_job, uint32_t total_items, double thr,
                   uint32_t* __restrict__ tile_counts, const uint32_t* __restrict__ tile_offs, uint32_t total_tiles,
                   double4* __restrict__ bin_lines, uint32_t refs_cap, int band_shift, int chunk_shift, Status* __restrict__ status,
                   Status* __restrict__ next_status) {
    // a single job travels in the kernel parameters (constant bank): no table upload, no dependent global load
    const JobDev* __restrict__ jobs = (n_jobs == 1 && !jobs_in) ? &one_job : jobs_in;
    // let the raster kernel behind us start its prologue as soon as every CTA of this grid is running (it waits for this
    // grid to complete before it reads anything we write)
    asm volatile("griddepcontrol.launch_dependents;");
    // the status block of the NEXT batch is cleared here (the two blocks alternate), so batches need no memset
    if (next_status && blockIdx.x == 0 && threadIdx.x < sizeof(Status) / 4) reinterpret_cast<uint32_t*>(next_status)[threadIdx.x] = 0u;
    // PASS 2 = single pass with fixed-capacity bins: refs_cap is the per-tile capacity
    if (PASS == 1) {
        if (status->nan_flag | status->depth_flag) return;
        const uint32_t n_refs = tile_offs[total_tiles];
        if (blockIdx.x == 0 && threadIdx.x == 0) status->n_refs = n_refs;
        if (n_refs > refs_cap) {
            if (blockIdx.x == 0 && threadIdx.x == 0) status->refs_overflow = 1u;
            return;
        }
    }
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    // PASS 2 runs on the packed grid (curve slots, then one thread per line item): `total_items` is its thread count
    const uint32_t total_slots = (PASS == 2) ? total_items : (total_items << DEPTH);
    SlotCtx c;
    uint32_t count = 0, n_refs = 0;
    auto bin_one = [&](const JobDev& job, double x0, double y0, double x1, double y1) {
        for_each_tile(job, x0, y0, x1, y1, band_shift, chunk_shift, [&](uint32_t key) {
            const uint32_t slot = atomicAdd(&tile_counts[key], 1u);
            if (PASS == 1) bin_lines[tile_offs[key] + slot] = make_double4(x0, y0, x1, y1);
            if (PASS == 2) {
                n_refs++;
                if (slot < refs_cap) {
                    bin_lines[(size_t)key * refs_cap + slot] = make_double4(x0, y0, x1, y1);
                } else {
                    status->refs_overflow = 1u;
                    atomicMax(&status->bin_max, slot + 1u);
                }
            }
        });
    };
    // PASS 2: the depth-first walks of a warp's lanes reach their leaves at different times; doing the tile arithmetic,
    // the counter atomic and the bin store inside the walk would make every lane sit through every other lane's
    // emission.  Leaves are parked in a per-warp shared queue instead (one shared atomic + a 32 B store) and binned after
    // the walk, 32 at a time with all lanes busy.
    __shared__ double4 q_line[PASS == 2 ? 4 : 1][PASS == 2 ? kWarpQueue : 1];
    __shared__ uint32_t q_job[PASS == 2 ? 4 : 1][PASS == 2 ? kWarpQueue : 1];
    __shared__ uint32_t q_n[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (PASS == 2) {
        if (lane == 0) q_n[warp] = 0;
        __syncwarp();
    }
    if (t < total_slots && (PASS == 2 ? slot_setup_packed<DEPTH>(jobs, n_jobs, t, thr, c, status) : slot_setup<DEPTH>(jobs, n_jobs, t, thr, c, status))) {
        const JobDev& job = jobs[c.job];
        auto emit = [&](double x0, double y0, double x1, double y1) {
            // band jobs: canvas rows -> band rows (JobDev::y_org; exact for every line that reaches the band; y - 0 == y otherwise)
            y0 = __dsub_rn(y0, job.y_org);
            y1 = __dsub_rn(y1, job.y_org);
            if (PASS == 2) {
                const uint32_t k = atomicAdd(&q_n[warp], 1u);
                if (k < (uint32_t)kWarpQueue) {
                    q_line[warp][k] = make_double4(x0, y0, x1, y1);
                    q_job[warp][k] = c.job;
                } else {
                    bin_one(job, x0, y0, x1, y1);  // queue full (a warp of deep curves): bin it here
                }
            } else {
                bin_one(job, x0, y0, x1, y1);
            }
        };
        // finite control points cannot produce NaN below: skip the per-node has_nans test (see seg_all_finite)
        count = seg_all_finite(c.seg, c.kind) ? slot_walk<false>(c, thr, status, emit) : slot_walk<true>(c, thr, status, emit);
    }
    if (PASS == 2) {
        // Binning, one lane per (line, band): a warp prefix sum over the band counts of 32 queued lines deals the
        // (line, band) pairs out evenly, so a long straight edge that crosses hundreds of bands (tv.path on the
        // 32768^2 canvas: up to 337) costs every lane a few dependent counter round trips instead of one lane hundreds
        // (round 1: that chain was 0.18 ms of the C5 step and 28 % of the C2 flatten kernel's stall samples).
        __syncwarp();
        const uint32_t n = min(q_n[warp], (uint32_t)kWarpQueue);
        auto bin_band = [&](const JobDev& job, const double4 l, int b) {
            for_band_tiles(job, l.x, l.y, l.z, l.w, b, band_shift, chunk_shift, [&](uint32_t key) {
                const uint32_t slot = atomicAdd(&tile_counts[key], 1u);
                n_refs++;
                if (slot < refs_cap) {
                    bin_lines[(size_t)key * refs_cap + slot] = l;
                } else {
                    status->refs_overflow = 1u;
                    atomicMax(&status->bin_max, slot + 1u);
                }
            });
        };
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t i = base + lane;
            int b0 = 0, span = 0;
            if (i < n) {
                const double4 l = q_line[warp][i];
                span = band_range(jobs[q_job[warp][i]], l.y, l.w, band_shift, b0);
            }
            int incl = span;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - span;
            for (int w0 = 0; w0 < total; w0 += 32) {
                const int w = w0 + lane;
                int src = 0;  // lanes whose pairs all come before pair w == the lane that owns it
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, incl, src + step - 1);
                    if (v <= w) src += step;
                }
                src = min(src, 31);
                const int sb0 = __shfl_sync(0xffffffffu, b0, src), sex = __shfl_sync(0xffffffffu, excl, src);
                if (w < total) bin_band(jobs[q_job[warp][base + src]], q_line[warp][base + src], sb0 + (w - sex));
            }
        }
    }
    if (PASS != 1) {  // total line count of the batch (statistics): one atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        if ((threadIdx.x & 31) == 0 && count) atomicAdd(&status->n_lines, count);
    }
    if (PASS == 2) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_refs += __shfl_xor_sync(0xffffffffu, n_refs, o);
        if ((threadIdx.x & 31) == 0 && n_refs) atomicAdd(&status->n_refs, n_refs);
    }
}

}  // namespace

void launch_flatten_count(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, uint32_t* slot_counts,
                          Status* status, cudaStream_t s) {
    uint32_t n = total_items * kSlotsPerItem + 1;
    flatten_kernel<false><<<(n + 127) / 128, 128, 0, s>>>(jobs, n_jobs, total_items, thr, slot_counts, nullptr, nullptr, 0, status);
}

void launch_flatten_emit(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, const uint32_t* slot_offs,
                         double4* lines, uint32_t lines_cap, Status* status, cudaStream_t s) {
    uint32_t n = total_items * kSlotsPerItem;
    if (n == 0) return;
    flatten_kernel<true><<<(n + 127) / 128, 128, 0, s>>>(jobs, n_jobs, total_items, thr, nullptr, slot_offs, lines, lines_cap, status);
}

static inline int log2i(int v) {
    int s = 0;
    while ((1 << s) < v) s++;
    return s;  // tile shapes are powers of two
}

void launch_flatten_bin_count(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, uint32_t* tile_counts,
                              int band_rows, int chunk_cols, Status* status, cudaStream_t s) {
    uint32_t n = total_items * kSlotsPerItem;
    if (n == 0) return;
    flatten_bin_kernel<0, kSlotDepth><<<(n + 127) / 128, 128, 0, s>>>(jobs, n_jobs, JobDev{}, total_items, thr, tile_counts, nullptr, 0, nullptr, 0,
                                                          log2i(band_rows), log2i(chunk_cols), status, nullptr);
}

void launch_flatten_bin_emit(const JobDev* jobs, uint32_t n_jobs, uint32_t total_items, double thr, const uint32_t* tile_offs,
                             uint32_t total_tiles, uint32_t* tile_cursor, double4* bin_lines, uint32_t refs_cap, int band_rows,
                             int chunk_cols, Status* status, cudaStream_t s) {
    uint32_t n = total_items * kSlotsPerItem;
    if (n == 0) return;
    flatten_bin_kernel<1, kSlotDepth><<<(n + 127) / 128, 128, 0, s>>>(jobs, n_jobs, JobDev{}, total_items, thr, tile_cursor, tile_offs, total_tiles,
                                                          bin_lines, refs_cap, log2i(band_rows), log2i(chunk_cols), status, nullptr);
}

int flatten_cut_depth(uint32_t total_items) {
    // Few items (a scene's fills, one stroked outline): cut every subdivision tree two levels deeper, 32 slots per curve —
    // four times the threads, each with a quarter of the subtree, because such batches are bound by the longest
    // depth-first walk, not by throughput.
    static const int forced = getenv("RGPU_CUT_DEPTH") ? atoi(getenv("RGPU_CUT_DEPTH")) : 0;  // tuning: 3..9
    if (forced >= 3 && forced <= 9) return forced;
    if (total_items <= kDeepestCutMaxItems) return kDeepestCutDepth;
    return total_items <= kDeepCutMaxItems ? kDeepCutDepth : kSlotDepth;
}

void launch_flatten_bin_fixed(const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t total_threads, int depth, double thr,
                              uint32_t* tile_counts, double4* bin_lines, uint32_t bin_cap, int band_rows, int chunk_cols, Status* status,
                              Status* next_status, cudaStream_t s) {
    if (total_threads == 0) return;
    const uint32_t grid = (total_threads + 127) / 128;
#define RGPU_FLAT(D)                                                                                                                      \
    flatten_bin_kernel<2, D><<<grid, 128, 0, s>>>(n_jobs == 1 ? nullptr : jobs, n_jobs, h_jobs[0], total_threads, thr, tile_counts, nullptr, 0, \
                                                  bin_lines, bin_cap, log2i(band_rows), log2i(chunk_cols), status, next_status)
    switch (depth) {
        case 4: RGPU_FLAT(4); break;
        case 5: RGPU_FLAT(5); break;
        case 6: RGPU_FLAT(6); break;
        case 7: RGPU_FLAT(7); break;
        case 8: RGPU_FLAT(8); break;
        case 9: RGPU_FLAT(9); break;
        default: RGPU_FLAT(kSlotDepth); break;
    }
#undef RGPU_FLAT
}

}  // namespace rgpu
