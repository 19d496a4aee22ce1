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

namespace rgpu {

namespace {

using namespace rs;

// Tile shapes: <CW columns, TH rows, THREADS>.  L = CW/32 columns per lane in the scan phase.
template <int CW, int TH, int THREADS>
struct TileCfg {
    static constexpr int kL = CW / 32;
    static constexpr int kPitch = CW + 4 * 32;               // 32 runs of L columns, 4 padding ints each
    static constexpr int kWarps = THREADS / 32;
    static constexpr int kRowBits = (TH <= 8) ? 3 : 6;
    static constexpr int kWarpSpanCap = (TH <= 8) ? 32 * TH : 512;  // per-warp span list entries
    static constexpr size_t smem_bytes() {
        return sizeof(int) * TH * kPitch + sizeof(double) * 4 * THREADS + sizeof(float) * THREADS +
               sizeof(unsigned short) * kWarpSpanCap * kWarps;
    }
};

template <int CW, int TH, int THREADS, bool EVENODD, class Cfg>
__device__ __forceinline__ void scan_rows(const JobDev& job, const PaintDev& s_paint, int* cells, const int* carry, const int* row_touched,
                                          int row0, int row1, int cx0, int mode, int tid) {
    constexpr int L = Cfg::kL;
    constexpr int NQ = L / 4;  // 128-bit words per lane run
    const int warp = tid >> 5, lane = tid & 31;
    const int wout = job.width_out;
    const int bw = min(wout - cx0, CW);  // visible columns of this tile
    for (int r = warp; r < row1 - row0; r += Cfg::kWarps) {
        const int acc = carry[r];
        int* bc = cells + r * Cfg::kPitch;
        const int y = row0 + r;
        if (row_touched[r]) {
            int v[L];
#pragma unroll
            for (int i = 0; i < NQ; i++) {
                const int4 q = *reinterpret_cast<const int4*>(bc + lane * (L + 4) + i * 4);
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
#pragma unroll
            for (int i = 0; i < NQ; i++) {
                const float4 cv = make_float4(coverage_from_fixed<EVENODD>(base + v[4 * i]), coverage_from_fixed<EVENODD>(base + v[4 * i + 1]),
                                              coverage_from_fixed<EVENODD>(base + v[4 * i + 2]), coverage_from_fixed<EVENODD>(base + v[4 * i + 3]));
                *reinterpret_cast<float4*>(bc + lane * (L + 4) + i * 4) = cv;
            }
        } else {
            float c = coverage_from_fixed<EVENODD>(acc);  // no line touched this row of the tile: constant coverage
            if (mode != kModeFill) {
                // straight from registers: no shared-memory round trip for empty rows (most of a sparse canvas)
                if (mode == kModeCoverage && c < 1e-6f) c = 0.f;
                const float4 cv = make_float4(c, c, c, c);
                float* out = reinterpret_cast<float*>(job.canvas) + job.origin + (unsigned long long)y * job.row_stride + cx0;
                const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
                for (int i = 0; i < CW / 128; i++) {
                    const int col = i * 128 + lane * 4;
                    if (col < bw) {
                        if (vec_ok && col + 3 < bw) {
                            __stcs(reinterpret_cast<float4*>(out + col), cv);
                        } else {
                            out[col] = c;
                            if (col + 1 < bw) out[col + 1] = c;
                            if (col + 2 < bw) out[col + 2] = c;
                            if (col + 3 < bw) out[col + 3] = c;
                        }
                    }
                }
                continue;
            }
            if (c < 1e-6f) continue;  // FILL: nothing to composite on this row
            const float4 cv = make_float4(c, c, c, c);
#pragma unroll
            for (int i = 0; i < NQ; i++) *reinterpret_cast<float4*>(bc + lane * (L + 4) + i * 4) = cv;
        }
        __syncwarp();
        // transposed read-back: lane l takes columns [128*i + 4l, +4): full 512 B coalesced 128-bit stores
        if (mode != kModeFill) {
            float* out = reinterpret_cast<float*>(job.canvas) + job.origin + (unsigned long long)y * job.row_stride + cx0;
            const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
            for (int i = 0; i < CW / 128; i++) {
                const int col = i * 128 + lane * 4;
                if (col < bw) {
                    float4 cv = *reinterpret_cast<const float4*>(bc + swz<L>(col));
                    if (mode == kModeCoverage) {  // mask_iter drops abs(alpha) < 1e-6 (src/rasterize.rs:348)
                        if (cv.x < 1e-6f) cv.x = 0.f;
                        if (cv.y < 1e-6f) cv.y = 0.f;
                        if (cv.z < 1e-6f) cv.z = 0.f;
                        if (cv.w < 1e-6f) cv.w = 0.f;
                    }
                    if (vec_ok && col + 3 < bw) {
                        __stcs(reinterpret_cast<float4*>(out + col), cv);  // streaming store: written once, never re-read
                    } else {
                        out[col] = cv.x;
                        if (col + 1 < bw) out[col + 1] = cv.y;
                        if (col + 2 < bw) out[col + 2] = cv.z;
                        if (col + 3 < bw) out[col + 3] = cv.w;
                    }
                }
            }
        } else {
            float4* out = reinterpret_cast<float4*>(job.canvas) + job.origin + (unsigned long long)y * job.row_stride + cx0;
            const float* covs = reinterpret_cast<const float*>(bc);
            for (int px = lane; px < bw; px += 32) {  // consecutive lanes composite consecutive pixels (16 B each)
                const float alpha = covs[swz<L>(px)];
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
        __syncwarp();
    }
}

// `one_job`: the launch covers a single job whose descriptor travels in the kernel parameters (constant bank:
// no dependent global loads before the tile can start).
template <int CW, int TH, int THREADS>
__global__ void __launch_bounds__(THREADS)
raster_kernel(const JobDev* __restrict__ jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first, const JobDev one_job,
              const PaintDev* __restrict__ paints, const uint32_t* __restrict__ tile_offs, uint32_t bin_cap,
              const double4* __restrict__ bin_lines, unsigned long long* __restrict__ tile_state, uint32_t epoch,
              uint32_t* __restrict__ ticket, const Status* __restrict__ status) {
    using Cfg = TileCfg<CW, TH, THREADS>;
    constexpr int L = Cfg::kL;
    static_assert(TH <= 64 && CW % 128 == 0 && L % 4 == 0, "tile shape");
    // dynamic shared memory: cells | piece constants | per-warp span lists
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* cells = reinterpret_cast<int*>(smem_raw);
    double* p_ax = reinterpret_cast<double*>(smem_raw + sizeof(int) * TH * Cfg::kPitch);
    double* p_ay = p_ax + THREADS;
    double* p_by = p_ay + THREADS;
    double* p_dxdy = p_by + THREADS;
    float* p_dir = reinterpret_cast<float*>(p_dxdy + THREADS);
    unsigned short* spans_all = reinterpret_cast<unsigned short*>(p_dir + THREADS);
    __shared__ int carry[TH];
    __shared__ int rowtot[TH];
    __shared__ int row_touched[TH];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_job;
    __shared__ PaintDev s_paint;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // dynamic tile id: a tile only ever waits (carry look-back) on tiles with smaller ids, which have started
    uint32_t my_ticket = 0;
    if (tid == 0) my_ticket = atomicAdd(ticket, 1u);
    const uint32_t bad = status->lines_overflow | status->refs_overflow | status->nan_flag | status->depth_flag;
    if (tid < TH) { carry[tid] = 0; rowtot[tid] = 0; row_touched[tid] = 0; }
    if (tid == 0) {
        const uint32_t t = tile_first + my_ticket;
        s_tile = t;
        s_job = (n_jobs == 1) ? job_first : job_first + find_job(n_jobs, t, [&](uint32_t k) { return jobs[job_first + k].tile_begin; });
    }
    __syncthreads();
    if (bad) return;
    const uint32_t tile = s_tile;
    const JobDev& job = (n_jobs == 1) ? one_job : jobs[s_job];
    const uint32_t lt = tile - job.tile_begin;
    const int band = (int)(lt / job.n_chunks);
    const int chunk = (int)(lt - (uint32_t)band * job.n_chunks);
    TileGeom g;
    g.row0 = band * TH;
    g.row1 = min(g.row0 + TH, job.height);
    g.cx0 = chunk * CW;
    g.wc = job.clamp_w;
    g.wci = (int)g.wc;
    g.tile_end = g.cx0 + min(CW, g.wci + 1 - g.cx0);  // columns that exist in the reference image (incl. overflow column)
    g.pitch = Cfg::kPitch;
    const int row0 = g.row0, row1 = g.row1, cx0 = g.cx0;
    const double wc = g.wc;
    const int mode = job.mode;

    if (mode == kModeFill && job.paint_index >= 0) {
        const int* src = reinterpret_cast<const int*>(&paints[job.paint_index]);
        int* dst = reinterpret_cast<int*>(&s_paint);
        for (int i = tid; i < (int)(sizeof(PaintDev) / 4); i += THREADS) dst[i] = src[i];
    }

    // ---- phase 1: accumulate the tile's lines.  Warps work independently: 32 lines per round per warp ------
    // (see warp_accumulate_round: 1a one line per lane, spans compacted per warp; 1b one lane per span)
    unsigned short* spans = spans_all + warp * Cfg::kWarpSpanCap;
    uint32_t rbeg, rend;
    if (bin_cap) {  // fixed-capacity bins: tile_offs holds the per-tile counts
        rbeg = tile * bin_cap;
        rend = rbeg + min(tile_offs[tile], bin_cap);
    } else {
        rbeg = tile_offs[tile];
        rend = tile_offs[tile + 1];
    }
    if (rbeg < rend) {  // tiles without lines never read their cells (row_touched stays 0): no need to clear them
        const int4 z = make_int4(0, 0, 0, 0);
        int4* c4 = reinterpret_cast<int4*>(cells);
        for (int i = tid; i < TH * Cfg::kPitch / 4; i += THREADS) c4[i] = z;
        __syncthreads();
    }
    for (uint32_t r0 = rbeg + warp * 32; r0 < rend; r0 += THREADS) {
        const uint32_t r = r0 + lane;
        const bool valid = r < rend;
        const double4 l = valid ? bin_lines[r] : make_double4(0, 0, 0, 0);
        warp_accumulate_round<L, Cfg::kRowBits, Cfg::kWarpSpanCap>(l, valid, g, cells, rowtot, row_touched, p_ax, p_ay, p_by, p_dxdy, p_dir,
                                                                  spans, tid);
    }
    __syncthreads();

    // ---- carry-in: decoupled look-back over the tiles to the left in this band --------------------------------
    // Every tile publishes its per-row totals (flag AGG), sums its predecessors' totals until it meets an
    // inclusive prefix, then publishes its own inclusive prefix.  Words carry the batch epoch, so the state needs
    // no clearing between batches.  Single-chunk jobs have no neighbours and skip all of this.
    if (job.n_chunks > 1) {
        if (tid < TH && tid < kStateRows) {
            const int agg = rowtot[tid];
            unsigned long long* st = tile_state + (size_t)tile * kStateRows + tid;
            const unsigned long long ep = (unsigned long long)epoch << 34;
            if (chunk == 0) {
                st_state(st, ep | kFlagPrefix | (unsigned long long)(uint32_t)agg);
            } else {
                st_state(st, ep | kFlagAgg | (unsigned long long)(uint32_t)agg);
                int sum = 0;
                for (int k = 1; k <= chunk; k++) {
                    const unsigned long long* ps = tile_state + (size_t)(tile - (uint32_t)k) * kStateRows + tid;
                    unsigned long long v;
                    do { v = ld_state(ps); } while ((uint32_t)(v >> 34) != epoch);
                    sum += (int)(uint32_t)v;
                    if ((v & (3ull << 32)) == kFlagPrefix) break;
                }
                carry[tid] = sum;
                st_state(st, ep | kFlagPrefix | (unsigned long long)(uint32_t)(sum + agg));
            }
        }
        __syncthreads();
    }

    // ---- phase 2: per-row scan, fill rule, store / composite --------------------------------------------
    // A warp takes a row; lane l owns L consecutive columns: serial prefix in registers, ONE warp scan of the 32
    // lane totals, coverage written back to shared memory in place (as floats) and read back transposed so that
    // global stores are full 512 B coalesced 128-bit accesses.
    if (job.rule == 1) scan_rows<CW, TH, THREADS, true, Cfg>(job, s_paint, cells, carry, row_touched, row0, row1, cx0, mode, tid);
    else scan_rows<CW, TH, THREADS, false, Cfg>(job, s_paint, cells, carry, row_touched, row0, row1, cx0, mode, tid);
}

// `From<LinColor> for RGBA`, src/color.rs:164-175 with the x86 l2s polynomial; `as u8` saturates
__device__ __forceinline__ unsigned char f2u8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (unsigned char)v;
}
__global__ void to_rgba8_kernel(const float4* __restrict__ lin, uchar4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 c = lin[i];
        float4 u = unmultiply(c);
        uchar4 o;
        o.x = f2u8(fadd(fmul(l2s_lane(u.x), 255.0f), 0.5f));
        o.y = f2u8(fadd(fmul(l2s_lane(u.y), 255.0f), 0.5f));
        o.z = f2u8(fadd(fmul(l2s_lane(u.z), 255.0f), 0.5f));
        o.w = f2u8(fadd(fmul(c.w, 255.0f), 0.5f));
        out[i] = o;
    }
}
__global__ void fill_color_kernel(float4* __restrict__ lin, size_t n, float4 color) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) lin[i] = color;
}
__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

}  // namespace


template <int CW, int TH, int THREADS>
static void launch_raster_t(const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first,
                            uint32_t n_tiles, const PaintDev* paints, const uint32_t* tile_offs, uint32_t bin_cap, const double4* bin_lines,
                            unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, const Status* status, cudaStream_t s) {
    constexpr size_t smem = TileCfg<CW, TH, THREADS>::smem_bytes();
    static bool configured[64] = {};  // per template instance and per device: the attribute belongs to the device's function
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(raster_kernel<CW, TH, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev] = true;
    }
    raster_kernel<CW, TH, THREADS><<<n_tiles, THREADS, smem, s>>>(jobs, n_jobs, job_first, tile_first, h_jobs[job_first], paints, tile_offs,
                                                                  bin_cap, bin_lines, tile_state, epoch, ticket, status);
}

TileShape raster_tile_shape(int variant) {
    switch (variant) {
        case 1: return TileShape{128, 64};   // canvases at most 128 px wide (larger than the fused small-canvas kernel takes)
        case 2: return TileShape{512, 8};    // tuning alternatives (RGPU_TILE_VARIANT), measured slower on C2 and C5
        case 3: return TileShape{512, 16};
        case 4: return TileShape{1024, 4};
        default: return TileShape{1024, 8};  // best of the r1 sweep: C2 46 us, C5 band 130 us (profiles/r1_tile_sweep.txt)
    }
}

void launch_raster(int variant, const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first,
                   uint32_t n_tiles, const PaintDev* paints, const uint32_t* tile_offs, uint32_t bin_cap, const double4* bin_lines,
                   unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, const Status* status, cudaStream_t s) {
    if (n_tiles == 0) return;
#define RGPU_LAUNCH(CW, TH, THREADS)                                                                                              \
    launch_raster_t<CW, TH, THREADS>(jobs, h_jobs, n_jobs, job_first, tile_first, n_tiles, paints, tile_offs, bin_cap, bin_lines, tile_state, \
                                     epoch, ticket, status, s)
    switch (variant) {
        case 1: RGPU_LAUNCH(128, 64, 256); break;
        case 2: RGPU_LAUNCH(512, 8, 128); break;
        case 3: RGPU_LAUNCH(512, 16, 128); break;
        case 4: RGPU_LAUNCH(1024, 4, 128); break;
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
