// K3 + K4 — signed-difference accumulation, per-row scan, fill rule and paint/composite, one CTA per
// (job, scanline band, column chunk) tile.
//
// Replaces, from the reference crate:
//   signed_difference_line      src/rasterize.rs:365-470  (+ split_at_zero_x :923-937)
//   signed_difference_to_mask   src/rasterize.rs:473-507
//   mask_iter's scan            src/rasterize.rs:333-353, FillRule::alpha_from_winding src/path.rs:32-46
//   Rasterizer::fill/fill_impl  src/rasterize.rs:70-115
//   Paint::at                   src/color.rs:357-360, src/grad.rs:24-30, 116-139, 202-211, 361-411,
//                               quadratic_solve src/utils.rs:205-231
//   LinColor maths              src/color.rs:308-354, s2l/l2s src/simd/x86.rs:197-244 (x86 polynomial variant)
//
// Determinism: the reference accumulates f64 deltas serially.  Here every (line,row) span is turned into the
// coverage of each pixel it crosses (the running sum of the reference's deltas), rounded to Q7.24 fixed point,
// and the DIFFERENCES of consecutive rounded coverages are added with integer shared-memory atomics.  Integer
// addition is associative, so the result does not depend on the order threads arrive, and the differences
// telescope: the cells of a span that fall in one tile sum to (rounded coverage at the tile's last column) -
// (rounded coverage left of its first column), so the per-row totals the tiles of a band exchange through the
// carry look-back are exact integers and the result is bit-identical from run to run.
#include "raster_device.cuh"

#include <type_traits>

namespace rgpu {

namespace {

using namespace rs;

// Tile shapes: <CW columns, TH rows, THREADS>.  L = CW/32 columns per lane in the scan phase.
template <int CW, int TH, int THREADS>
struct TileCfg {
    static constexpr int kL = CW / 32;
    static constexpr int kWarps = THREADS / 32;
    static constexpr int kRowBits = (TH <= 8) ? 3 : 6;
    // per-warp span list: (source lane, band row) packed in a byte when the row fits 3 bits
    using SpanT = typename std::conditional<(TH <= 8), unsigned char, unsigned short>::type;
    static constexpr int kWarpSpanCap = (TH <= 8) ? 208 : 512;  // lanes that do not fit do their rows serially
    static constexpr size_t kCellBytes = sizeof(int) * TH * CW;
    static constexpr size_t kPieceBytes = sizeof(double) * 4 * THREADS;  // reused for the paint in the scan phase
    static constexpr size_t smem_bytes() { return kCellBytes + kPieceBytes + sizeof(SpanT) * kWarpSpanCap * kWarps; }
    static_assert(kPieceBytes >= sizeof(PaintDev), "the paint is staged over the piece constants");
};

// Phase 2 for the rows of one warp.  Lane l owns L consecutive columns: serial prefix in registers, ONE warp scan of the
// 32 lane totals, coverage written back to shared memory in place (as floats) and read back transposed so that global
// stores are full 512 B coalesced 128-bit accesses.  Rows no line touched are written straight from registers.
template <int CW, int TH, int THREADS, bool EVENODD, bool FILL, class Cfg>
__device__ __forceinline__ void scan_rows(const JobDev& job, const PaintDev& s_paint, int* cells, const int* carry, const int* row_touched,
                                          int* row_live, int row0, int row1, int cx0, int mode_in, int tid, const Fix fix, const uint32_t n_lines,
                                          Status* __restrict__ status) {
    // FILL = false: the launch holds no FILL job and the paint / composite code is not even compiled in
    const int mode = (!FILL && mode_in == kModeFill) ? kModeMask : mode_in;
    constexpr int L = Cfg::kL;
    constexpr int NQ = L / 4;  // 128-bit words per lane run
    const int warp = tid >> 5, lane = tid & 31;
    const int bw = min(job.width_out - cx0, CW);  // visible columns of this tile
    for (int r = warp; r < row1 - row0; r += Cfg::kWarps) {
        const int acc = carry[r];
        int* bc = cells + r * CW;
        const int y = row0 + r;
        // winding guard (NonZero): inside the tile a row's winding moves away from its carry-in by at most one per line of
        // the tile, so only dense tiles or large carries look at their pixels
        const bool wcheck = !EVENODD && (abs(acc) >> job.fix_shift) + (int)n_lines >= (fix.guard >> job.fix_shift);
        if (wcheck && lane == 0 && abs(acc) >= fix.guard) status->winding_flag = 1u;
        if (row_touched[r]) {
            int v[L];
#pragma unroll
            for (int i = 0; i < NQ; i++) {
                const int4 q = *reinterpret_cast<const int4*>(bc + swz<true>(lane * L + i * 4));
                v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
            }
#pragma unroll
            for (int i = 1; i < L; i++) v[i] += v[i - 1];
            int incl = v[L - 1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int nb = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += nb;
            }
            const int base = acc + incl - v[L - 1];
            if (wcheck) {
                bool risk = false;
#pragma unroll
                for (int i = 0; i < NQ; i++) risk = risk || winding_risk(base + v[4 * i], base + v[4 * i + 1], base + v[4 * i + 2], base + v[4 * i + 3], fix);
                if (risk) status->winding_flag = 1u;
            }
#pragma unroll
            for (int i = 0; i < NQ; i++) {
                const float4 cv = make_float4(coverage_from_fixed<EVENODD>(base + v[4 * i], fix), coverage_from_fixed<EVENODD>(base + v[4 * i + 1], fix),
                                              coverage_from_fixed<EVENODD>(base + v[4 * i + 2], fix), coverage_from_fixed<EVENODD>(base + v[4 * i + 3], fix));
                *reinterpret_cast<float4*>(bc + swz<true>(lane * L + i * 4)) = cv;
            }
        } else {
            float c = coverage_from_fixed<EVENODD>(acc, fix);  // no line touched this row of the tile: constant coverage
            if (!FILL || mode != kModeFill) {
                // straight from registers: no shared-memory round trip for empty rows (most of a sparse canvas)
                if (mode == kModeCoverage && c < 1e-6f) c = 0.f;
                const float4 cv = make_float4(c, c, c, c);
                float* out = reinterpret_cast<float*>(job.canvas) + job.origin + (unsigned long long)y * job.row_stride + cx0;
                if (bw == CW && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
#pragma unroll
                    for (int i = 0; i < CW / 128; i++) __stcs(reinterpret_cast<float4*>(out + i * 128 + lane * 4), cv);
                } else {
                    for (int col = lane; col < bw; col += 32) out[col] = c;
                }
                continue;
            }
            // FILL: stage the constant coverage like a scanned row (or mark the row as having nothing to composite)
            if (lane == 0) row_live[r] = c >= 1e-6f;
            if (c < 1e-6f) continue;
            const float4 cv = make_float4(c, c, c, c);
#pragma unroll
            for (int i = 0; i < NQ; i++) *reinterpret_cast<float4*>(bc + swz<true>(lane * L + i * 4)) = cv;
            continue;
        }
        if (FILL && mode == kModeFill) {
            if (lane == 0) row_live[r] = 1;
            continue;  // composited below by the whole CTA
        }
        __syncwarp();
        // transposed read-back: lane l takes columns [128*i + 4l, +4): full 512 B coalesced 128-bit stores
        float* out = reinterpret_cast<float*>(job.canvas) + job.origin + (unsigned long long)y * job.row_stride + cx0;
        if (bw == CW && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {  // whole, aligned tile row: no per-word tests
#pragma unroll
            for (int i = 0; i < CW / 128; i++) {
                const int col = i * 128 + lane * 4;
                float4 cv = *reinterpret_cast<const float4*>(bc + swz<true>(col));
                if (mode == kModeCoverage) {  // mask_iter drops abs(alpha) < 1e-6 (src/rasterize.rs:348)
                    if (cv.x < 1e-6f) cv.x = 0.f;
                    if (cv.y < 1e-6f) cv.y = 0.f;
                    if (cv.z < 1e-6f) cv.z = 0.f;
                    if (cv.w < 1e-6f) cv.w = 0.f;
                }
                __stcs(reinterpret_cast<float4*>(out + col), cv);  // streaming store: written once, never re-read
            }
        } else {
            const float* covs = reinterpret_cast<const float*>(bc);
            for (int col = lane; col < bw; col += 32) {
                float cvx = covs[swz<true>(col)];
                if (mode == kModeCoverage && cvx < 1e-6f) cvx = 0.f;
                out[col] = cvx;
            }
        }
        __syncwarp();
    }
    if (!FILL || mode != kModeFill) return;
    // ---- K4: paint + composite.  The per-pixel paint evaluation (f64 point transform, gradient offset, stop search,
    // sRGB -> linear) is hundreds of dependent instructions, so the tile's pixels are spread over ALL threads of the CTA:
    // consecutive threads take consecutive pixels of a row (16 B each, coalesced read-modify-write).
    __syncthreads();
    const int rows = row1 - row0;
    const float* covs = reinterpret_cast<const float*>(cells);
    constexpr int TPR = (THREADS >= TH) ? THREADS / TH : 1;  // threads per tile row (power of two)
    for (int r = tid / TPR; r < rows; r += (THREADS / TPR)) {
        if (!row_live[r]) continue;
        const int y = row0 + r;
        float4* out = reinterpret_cast<float4*>(job.canvas) + job.origin + (unsigned long long)y * job.row_stride + cx0;
        for (int px = tid % TPR; px < bw; px += TPR) {
            const float alpha = covs[r * CW + swz<true>(px)];
            if (alpha >= 1e-6f) {
                float4 color = (job.paint_index >= 0) ? paint_at(s_paint, cx0 + px, y) : make_float4(0.f, 0.f, 0.f, 0.f);
                // with_alpha: self * (alpha as f32), src/color.rs:347-349
                color = make_float4(fmul(color.x, alpha), fmul(color.y, alpha), fmul(color.z, alpha), fmul(color.w, alpha));
                // blend_over: other + self * (1 - other.alpha), src/color.rs:342-344
                float4 dstc = out[px];
                const float k = fsub(1.0f, color.w);
                dstc = make_float4(fadd(color.x, fmul(dstc.x, k)), fadd(color.y, fmul(dstc.y, k)), fadd(color.z, fmul(dstc.z, k)),
                                   fadd(color.w, fmul(dstc.w, k)));
                out[px] = dstc;
            }
        }
    }
}

// `one_job`: the launch covers a single job whose descriptor travels in the kernel parameters (constant bank:
// no dependent global loads before the tile can start).
template <int CW, int TH, int THREADS, bool FILL>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 6 : (THREADS >= 512 ? 2 : 1)))
raster_kernel(const JobDev* __restrict__ jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first, const JobDev one_job,
              const PaintDev* __restrict__ paints, uint32_t* __restrict__ tile_offs, uint32_t bin_cap,
              const double4* __restrict__ bin_lines, unsigned long long* __restrict__ tile_state, uint32_t epoch,
              uint32_t* __restrict__ ticket, Status* status, uint32_t zero_early, uint32_t n_tiles, uint32_t late_wait) {
    using Cfg = TileCfg<CW, TH, THREADS>;
    using SpanT = typename Cfg::SpanT;
    static_assert(TH <= 64 && CW % 128 == 0 && Cfg::kL % 4 == 0, "tile shape");
    // dynamic shared memory: cells | piece constants (later: the paint) | per-warp span lists
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* cells = reinterpret_cast<int*>(smem_raw);
    double* p_ax = reinterpret_cast<double*>(smem_raw + Cfg::kCellBytes);
    double* p_ay = p_ax + THREADS;
    double* p_by = p_ay + THREADS;
    double* p_dxdy = p_by + THREADS;
    SpanT* spans_all = reinterpret_cast<SpanT*>(p_dxdy + THREADS);
    const PaintDev& s_paint = *reinterpret_cast<const PaintDev*>(p_ax);
    __shared__ int carry[TH];
    __shared__ int rowtot[TH];
    __shared__ int row_touched[TH];
    __shared__ uint32_t s_tile, s_job, s_chunk, s_band, s_count, s_bad;
    __shared__ unsigned long long s_early[TH <= 8 ? TH : 1];  // left neighbour's look-back words, probed by thread 0 (below)

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // dynamic tile id: a tile only ever waits (carry look-back) on tiles with smaller ids, which have started.
    // (A persistent variant — one CTA per resident slot looping over tickets, next ticket prefetched — was measured
    // slower on every workload: C2 raster 39 -> 43 us, C5 125 -> 127 us; the hardware CTA scheduler refills slots faster
    // than a loop with two extra barriers per tile.)
    uint32_t my_ticket = 0;
    if (tid == 0) my_ticket = atomicAdd(ticket, 1u);
    // dense batches clear the cells while the ticket is on its way; sparse ones (most tiles without a line) clear only
    // the tiles that need it, once the tile is known
    auto clear_cells = [&]() {
        const int4 z = make_int4(0, 0, 0, 0);
        int4* c4 = reinterpret_cast<int4*>(cells);
#pragma unroll 4
        for (int i = tid; i < TH * CW / 4; i += THREADS) c4[i] = z;
    };
    if (zero_early) clear_cells();
    if (tid < TH) { carry[tid] = 0; rowtot[tid] = 0; row_touched[tid] = 0; }
    // Programmatic dependent launch: everything above overlaps the tail of the flatten kernel; the bins, their counters and
    // the status flags it writes are read only after it has completed (no-op when launched without the attribute).
    // `late_wait` (the 2nd, 3rd ... launch of an ordered batch): the kernel before us is the previous fill of the same
    // canvas; only the composite at the end depends on it, so the wait moves there and this tile's accumulation overlaps
    // the previous fill's compositing.  The flatten kernel is long complete by then: the first raster launch lets its
    // dependents go only after its own wait has returned.
    if (!late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    // volatile: loads through a `const __restrict__` pointer are invariant to the compiler and may be hoisted above the
    // wait, where the flatten kernel has not raised its flags yet
    // (thread 0 alone reads them, next to its ticket, and broadcasts the verdict with the tile)
    const volatile uint32_t* flags = reinterpret_cast<const volatile uint32_t*>(status);
    // Tiles are taken in CHUNK-major order (all bands of column chunk 0, then chunk 1, ...): the left neighbour a tile's
    // carry depends on was started a whole column of bands earlier, so it has normally published its inclusive prefix by
    // the time this tile looks back — band-major order would start the tiles of a band together and make every tile wait
    // for the slowest one to its left.  Bins and look-back state stay indexed band-major.
    if (tid == 0) {
        s_bad = flags[0] | flags[1] | flags[2] | flags[3];  // nan, depth, lines_overflow, refs_overflow
        if (my_ticket == n_tiles - 1) *ticket = 0u;  // every ticket of this launch is drawn: leave the counter clean
        const uint32_t t = tile_first + my_ticket;
        const uint32_t j = (n_jobs == 1) ? job_first : job_first + find_job(n_jobs, t, [&](uint32_t k) { return jobs[job_first + k].tile_begin; });
        const JobDev& jb = (n_jobs == 1) ? one_job : jobs[j];
        const uint32_t lt = t - jb.tile_begin;
        const uint32_t chunk = lt / jb.n_bands;
        const uint32_t band = lt - chunk * jb.n_bands;
        const uint32_t tile = jb.tile_begin + band * jb.n_chunks + chunk;
        s_job = j;
        s_chunk = chunk;
        s_band = band;
        s_tile = tile;
        // Early look-back probe: in chunk-major order the left neighbour has normally published its inclusive prefixes
        // long ago.  Their state words are requested here, together with the bin counter, so that one memory round trip
        // serves both; they are consumed after phase 1.
        unsigned long long ev[TH <= 8 ? TH : 1];
        const bool probe = TH <= 8 && jb.n_chunks > 1 && chunk > 0;
        if (probe) {
#pragma unroll
            for (int r = 0; r < (TH <= 8 ? TH : 1); r++) ev[r] = ld_state(tile_state + (size_t)(tile - 1u) * kStateRows + r);
        }
        if (bin_cap) {  // fixed-capacity bins: tile_offs holds the per-tile counts; leave the counter clean for the next batch
            s_count = min(tile_offs[tile], bin_cap);
            tile_offs[tile] = 0u;
        }
#pragma unroll
        for (int r = 0; r < (TH <= 8 ? TH : 1); r++) s_early[r] = probe ? ev[r] : 0ull;
    }
    __syncthreads();
    if (s_bad) return;
    const JobDev& job = (n_jobs == 1) ? one_job : jobs[s_job];
    const uint32_t tile = s_tile;
    const int chunk = (int)s_chunk, band = (int)s_band;
    TileGeom g;
    g.row0 = band * TH;
    g.row1 = min(g.row0 + TH, job.height);
    g.cx0 = chunk * CW;
    g.wc = job.clamp_w;
    g.wci = (int)g.wc;
    g.tile_end = g.cx0 + min(CW, g.wci + 1 - g.cx0);  // columns that exist in the reference image (incl. overflow column)
    g.pitch = CW;
    const Fix fix = make_fix(job.fix_shift);
    g.fix_scale = fix.scale;
    const int row0 = g.row0, row1 = g.row1, cx0 = g.cx0;
    const int mode = job.mode;
    constexpr int LPR = (THREADS / TH >= 32) ? 32 : (THREADS / TH);  // lanes per row in the look-back (16 for 1024 x 8 / 128)
    static_assert(LPR >= 1 && (LPR & (LPR - 1)) == 0, "lanes per row must be a power of two");

    // ---- phase 1: accumulate the tile's lines.  The lines are split evenly over the warps, which then work
    // independently, 32 lines per round (see warp_accumulate_round: 1a one line per lane, spans compacted per warp;
    // 1b one lane per span) ------------------------------------------------------------------------------------------
    SpanT* spans = spans_all + warp * Cfg::kWarpSpanCap;
    uint32_t rbeg, rend;
    if (bin_cap) {
        rbeg = tile * bin_cap;
        rend = rbeg + s_count;
    } else {
        rbeg = tile_offs[tile];
        rend = tile_offs[tile + 1];
    }
    if (!zero_early && rbeg < rend) {  // tiles without lines never read their cells (row_touched stays 0)
        clear_cells();
        __syncthreads();
    }
    {
        const uint32_t per = (rend - rbeg + Cfg::kWarps - 1) / Cfg::kWarps;
        const uint32_t wbeg = rbeg + (uint32_t)warp * per, wend = min(wbeg + per, rend);
        for (uint32_t r0 = wbeg; r0 < wend; r0 += 32) {
            const uint32_t r = r0 + lane;
            const bool valid = r < wend;
            const double4 l = valid ? bin_lines[r] : make_double4(0, 0, 0, 0);
            warp_accumulate_round<true, Cfg::kRowBits, Cfg::kWarpSpanCap, SpanT>(l, valid, g, cells, rowtot, row_touched, p_ax, p_ay, p_by, p_dxdy,
                                                                           spans, tid);
        }
    }
    __syncthreads();
    // the piece constants are dead: stage the paint over them
    if (FILL && mode == kModeFill && job.paint_index >= 0) {
        const int* src = reinterpret_cast<const int*>(&paints[job.paint_index]);
        int* dst = reinterpret_cast<int*>(p_ax);
        for (int i = tid; i < (int)(sizeof(PaintDev) / 4); i += THREADS) dst[i] = src[i];
    }

    // ---- carry-in: decoupled look-back over the tiles to the left in this band --------------------------------
    // Every tile publishes its per-row totals (flag AGG), sums its predecessors' totals until it meets an
    // inclusive prefix, then publishes its own inclusive prefix.  All rows look back at once, each with a window of
    // predecessors.  Words carry the batch epoch, so the state needs no clearing between batches.
    // Single-chunk jobs have no neighbours and skip all of this.
    if (job.n_chunks > 1) {
        // every row of the tile looks back at once: a group of LPR lanes per row, lane k of the group on predecessor k
        const unsigned long long ep = (unsigned long long)epoch << 34;
        const int r = tid / LPR, gl = tid % LPR;
        if (r < TH && r < kStateRows) {  // uniform per warp: a warp holds 32 / LPR whole rows
            const int agg = rowtot[r];
            unsigned long long* st = tile_state + (size_t)tile * kStateRows + r;
            if (chunk == 0) {
                if (gl == 0) st_state(st, ep | kFlagPrefix | (unsigned long long)(uint32_t)agg);
            } else {
                if (gl == 0) st_state(st, ep | kFlagAgg | (unsigned long long)(uint32_t)agg);
                const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((lane / LPR) * LPR));
                int sum = 0;
                // the early probe settles the row when it already saw this batch's inclusive prefix
                const unsigned long long early = (TH <= 8) ? s_early[r] : 0ull;
                bool done = (uint32_t)(early >> 34) == epoch && (early & (3ull << 32)) == kFlagPrefix;
                if (done) sum = (int)(uint32_t)early;
                for (int k0 = 1; k0 <= chunk && !__all_sync(0xffffffffu, done); k0 += LPR) {  // same trip count for every row of the tile
                    const int k = k0 + gl;
                    const bool look = !done && k <= chunk;
                    unsigned long long v = 0;
                    if (look) {
                        const unsigned long long* ps = tile_state + (size_t)(tile - (uint32_t)k) * kStateRows + r;
                        do { v = ld_state(ps); } while ((uint32_t)(v >> 34) != epoch);
                    }
                    const unsigned pm = __ballot_sync(0xffffffffu, look && (v & (3ull << 32)) == kFlagPrefix) & gmask;
                    // nearest predecessor of this row that already holds an inclusive prefix
                    const int first = pm ? (__ffs(pm) - 1) % LPR : LPR - 1;
                    int contrib = (look && gl <= first) ? (int)(uint32_t)v : 0;
#pragma unroll
                    for (int o = LPR / 2; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                    sum += contrib;
                    done = done || pm != 0;
                    if (__all_sync(0xffffffffu, done)) break;
                }
                if (gl == 0) {
                    carry[r] = sum;
                    st_state(st, ep | kFlagPrefix | (unsigned long long)(uint32_t)(sum + agg));
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: per-row scan, fill rule, store / composite (rowtot is dead: FILL reuses it as a per-row live flag) ----
    if (late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous fill of this canvas is complete
    if (job.rule == 1) scan_rows<CW, TH, THREADS, true, FILL, Cfg>(job, s_paint, cells, carry, row_touched, rowtot, row0, row1, cx0, mode, tid, fix, rend - rbeg, status);
    else scan_rows<CW, TH, THREADS, false, FILL, Cfg>(job, s_paint, cells, carry, row_touched, rowtot, row0, row1, cx0, mode, tid, fix, rend - rbeg, status);
}

__global__ void to_rgba8_kernel(const float4* __restrict__ lin, uchar4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        out[i] = lin_to_rgba8(lin[i]);
    }
}
__global__ void fill_color_kernel(float4* __restrict__ lin, size_t n, float4 color) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) lin[i] = color;
}
__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

}  // namespace


template <int CW, int TH, int THREADS, bool FILL>
static void launch_raster_t(const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first,
                            uint32_t n_tiles, const PaintDev* paints, uint32_t* tile_offs, uint32_t bin_cap, const double4* bin_lines,
                            unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, Status* status, bool zero_early, int pdl, cudaStream_t s) {
    constexpr size_t smem = TileCfg<CW, TH, THREADS>::smem_bytes();
    static bool configured[64] = {};  // per template instance and per device: the attribute belongs to the device's function
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(raster_kernel<CW, TH, THREADS, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(raster_kernel<CW, TH, THREADS, FILL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured[dev] = true;
    }
    // `pdl`: launched with programmatic stream serialization — the CTAs may start while the preceding kernel of the stream
    // (the flatten pass, which triggers early) is still draining, and block in griddepcontrol.wait until it has completed
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;  // pdl: 0 = plain launch, 1 = wait at the top (behind the flatten kernel), 2 = late wait
    cudaLaunchKernelEx(&cfg, raster_kernel<CW, TH, THREADS, FILL>, jobs, n_jobs, job_first, tile_first, h_jobs[job_first], paints, tile_offs, bin_cap,
                       bin_lines, tile_state, epoch, ticket, status, zero_early ? 1u : 0u, n_tiles, pdl == 2 ? 1u : 0u);
}

TileShape raster_tile_shape(int variant) {
    switch (variant) {
        case 1: return TileShape{128, 64};   // canvases at most 128 px wide (larger than the fused small-canvas kernel takes)
        case 2: return TileShape{1024, 8};   // same tile, 512 threads: FILL launches (per-pixel paint evaluation dominates)
        default: return TileShape{1024, 8};  // best of the r1 sweep over {256,512,1024} x {4,8,16} (profiles/r1_tile_sweep.txt)
    }
}

void launch_raster(int variant, const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first,
                   uint32_t n_tiles, const PaintDev* paints, uint32_t* tile_offs, uint32_t bin_cap, const double4* bin_lines,
                   unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, Status* status, bool zero_early, int pdl, cudaStream_t s) {
    if (n_tiles == 0) return;
    // launches without a FILL job run the instantiation that carries no paint / composite code (fewer registers)
    bool fill = false;
    for (uint32_t j = 0; j < n_jobs && !fill; j++) fill = h_jobs[job_first + j].mode == kModeFill;
#define RGPU_LAUNCH(CW, TH, THREADS)                                                                                                  \
    do {                                                                                                                              \
        if (fill)                                                                                                                     \
            launch_raster_t<CW, TH, THREADS, true>(jobs, h_jobs, n_jobs, job_first, tile_first, n_tiles, paints, tile_offs, bin_cap, bin_lines, \
                                                   tile_state, epoch, ticket, status, zero_early, pdl, s);                            \
        else                                                                                                                          \
            launch_raster_t<CW, TH, THREADS, false>(jobs, h_jobs, n_jobs, job_first, tile_first, n_tiles, paints, tile_offs, bin_cap, bin_lines, \
                                                    tile_state, epoch, ticket, status, zero_early, pdl, s);                           \
    } while (0)
    switch (variant) {
        case 1: RGPU_LAUNCH(128, 64, 256); break;
        case 2: RGPU_LAUNCH(1024, 8, 512); break;
        default: RGPU_LAUNCH(1024, 8, 128); break;
    }
#undef RGPU_LAUNCH
}

void launch_to_rgba8(const float4* lin, uchar4* out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    int grid = (int)min((size_t)148 * 16, (n + 255) / 256);
    to_rgba8_kernel<<<grid, 256, 0, s>>>(lin, out, n);
}
void launch_fill_color(float4* lin, size_t n, float4 color, cudaStream_t s) {
    if (n == 0) return;
    int grid = (int)min((size_t)148 * 16, (n + 255) / 256);
    fill_color_kernel<<<grid, 256, 0, s>>>(lin, n, color);
}
void launch_f32_to_f64(const float* in, double* out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    int grid = (int)min((size_t)148 * 16, (n + 255) / 256);
    f32_to_f64_kernel<<<grid, 256, 0, s>>>(in, out, n);
}

}  // namespace rgpu
