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
// telescope: whatever subset of a span's cells falls left of a tile sums to exactly the rounded coverage at the
// tile edge, so tiles of one band agree on their carry-in without communicating.
#include "rgpu_internal.cuh"

namespace rgpu {

namespace {

constexpr double kEps = 2.220446049250313e-16;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- colour maths (f32, never contracted: the reference uses plain SSE mul/add) -------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// src/simd/x86.rs:217-244, one lane
__device__ __forceinline__ float s2l_lane(float v) {
    float x1 = fsub(fmul(2.0843103538116825f, v), 1.0843103538116827f);
    float x2 = fmul(x1, x1);
    float x3 = fmul(x2, x1);
    float high = fadd(fadd(fadd(0.23361048543711943f, fmul(0.4665843122387033f, x1)), fmul(0.26901741378006355f, x2)),
                      fmul(0.031661580753065945f, x3));
    return (v <= 0.04045f) ? fmul(v, 0.07739938080495357f) : high;
}
// src/simd/x86.rs:197-214, one lane
__device__ __forceinline__ float l2s_lane(float x0) {
    float x1 = __fsqrt_rn(x0);
    float x2 = __fsqrt_rn(x1);
    float x3 = __fsqrt_rn(x2);
    float high = fsub(fadd(fadd(fmul(-0.01848558f, x0), fmul(0.6445592f, x1)), fmul(0.70994765f, x2)), fmul(0.33605254f, x3));
    return (x0 <= 0.0031308f) ? fmul(x0, 12.92f) : high;
}
// LinColor::unmultiply, src/color.rs:308-317
__device__ __forceinline__ float4 unmultiply(float4 c) {
    if (c.w <= 1e-6f) return make_float4(0.f, 0.f, 0.f, 0.f);
    return make_float4(__fdiv_rn(c.x, c.w), __fdiv_rn(c.y, c.w), __fdiv_rn(c.z, c.w), __fdiv_rn(c.w, c.w));
}
// LinColor::into_linear, src/color.rs:330-332 (all four lanes go through the polynomial, alpha included)
__device__ __forceinline__ float4 into_linear(float4 c) {
    float4 u = unmultiply(c);
    float a = c.w;
    return make_float4(fmul(s2l_lane(u.x), a), fmul(s2l_lane(u.y), a), fmul(s2l_lane(u.z), a), fmul(s2l_lane(u.w), a));
}

// f64::rem_euclid
__device__ __forceinline__ double rem_euclid(double x, double rhs) {
    double r = fmod(x, rhs);
    return r < 0.0 ? r + fabs(rhs) : r;
}

// GradStops::at, src/grad.rs:116-139
__device__ float4 stops_at(const PaintDev& P, double t) {
    int lo = 0, hi = P.n_stops;
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (P.stop_pos[mid] < t) lo = mid + 1; else hi = mid;
    }
    int index = lo, size = P.n_stops;
    if (index == 0) return make_float4(P.stop_col[0][0], P.stop_col[0][1], P.stop_col[0][2], P.stop_col[0][3]);
    if (index == size)
        return make_float4(P.stop_col[size - 1][0], P.stop_col[size - 1][1], P.stop_col[size - 1][2], P.stop_col[size - 1][3]);
    double pos0 = P.stop_pos[index - 1], pos1 = P.stop_pos[index];
    float r = (float)((t - pos0) / (pos1 - pos0));
    float ir = fsub(1.0f, r);
    const float* c0 = P.stop_col[index - 1];
    const float* c1 = P.stop_col[index];
    // lerp: other * t + self * (1 - t), src/color.rs:352-354
    return make_float4(fadd(fmul(c1[0], r), fmul(c0[0], ir)), fadd(fmul(c1[1], r), fmul(c0[1], ir)),
                       fadd(fmul(c1[2], r), fmul(c0[2], ir)), fadd(fmul(c1[3], r), fmul(c0[3], ir)));
}

// utils::quadratic_solve + GradRadial::offset root selection, src/utils.rs:205-231, src/grad.rs:361-396
__device__ bool radial_offset(const PaintDev& P, double px, double py, double& out) {
    double cdx = __dsub_rn(P.p0x, P.p1x), cdy = __dsub_rn(P.p0y, P.p1y);
    double pdx = __dsub_rn(px, P.p1x), pdy = __dsub_rn(py, P.p1y);
    double rd = __dsub_rn(P.r0, P.r1);
    double a = __dsub_rn(__dadd_rn(__dmul_rn(cdx, cdx), __dmul_rn(cdy, cdy)), __dmul_rn(rd, rd));
    double b = __dmul_rn(-2.0, __dadd_rn(__dadd_rn(__dmul_rn(cdx, pdx), __dmul_rn(cdy, pdy)), __dmul_rn(P.r1, rd)));
    double c = __dsub_rn(__dadd_rn(__dmul_rn(pdx, pdx), __dmul_rn(pdy, pdy)), __dmul_rn(P.r1, P.r1));
    if (fabs(a) < kEps) {
        if (fabs(b) > kEps) { out = __ddiv_rn(-c, b); return true; }
        return false;
    }
    double disc = __dsub_rn(__dmul_rn(b, b), __dmul_rn(__dmul_rn(4.0, a), c));
    if (fabs(disc) < kEps) { out = __ddiv_rn(-b, __dmul_rn(2.0, a)); return true; }
    if (disc > 0.0) {
        double sq = __dsqrt_rn(disc);
        double t0, t1;
        if (b >= 0.0) {
            double mul = __dsub_rn(-b, sq);
            t0 = __ddiv_rn(mul, __dmul_rn(2.0, a));
            t1 = __ddiv_rn(__dmul_rn(2.0, c), mul);
        } else {
            double mul = __dadd_rn(-b, sq);
            t0 = __ddiv_rn(__dmul_rn(2.0, c), mul);
            t1 = __ddiv_rn(mul, __dmul_rn(2.0, a));
        }
        out = isnan(t0) ? t1 : (isnan(t1) ? t0 : fmax(t0, t1));
        return true;
    }
    return false;
}

// Paint::at for a pixel centre, after pixel_tr (src/rasterize.rs:93-96)
__device__ float4 paint_at(const PaintDev& P, int x, int y) {
    if (P.kind == 0) return make_float4(P.solid[0], P.solid[1], P.solid[2], P.solid[3]);
    double fx = (double)x + 0.5, fy = (double)y + 0.5;
    const double* m = P.pixel_tr;
    double px = __dadd_rn(__dadd_rn(__dmul_rn(fx, m[0]), __dmul_rn(fy, m[1])), m[2]);
    double py = __dadd_rn(__dadd_rn(__dmul_rn(fx, m[3]), __dmul_rn(fy, m[4])), m[5]);
    double t;
    if (P.kind == 1) {
        // (point - start).dot(dir), src/grad.rs:204
        t = __dadd_rn(__dmul_rn(__dsub_rn(px, P.p0x), P.dirx), __dmul_rn(__dsub_rn(py, P.p0y), P.diry));
    } else {
        if (!radial_offset(P, px, py, t)) return make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (P.spread == 1) t = rem_euclid(t, 1.0);
    else if (P.spread == 2) t = fabs(rem_euclid(t + 1.0, 2.0) - 1.0);
    float4 c = stops_at(P, t);
    return P.linear_colors ? c : into_linear(c);
}

// ---- coverage from the fixed-point winding ---------------------------------------------------------------
// NonZero: min(|w|, 1) (the reference's `value < 1e-6 -> 0` only matters to mask_iter's pixel dropping, which
// the COVERAGE / FILL paths apply themselves; below 1e-6 the two differ by < 1e-6).  EvenOdd is exact in integers.
__device__ __forceinline__ float coverage_from_fixed(int acc, int rule) {
    constexpr float kInv = 1.0f / 16777216.0f;
    if (rule == 1) {
        // abs(((w + 1) rem_euclid 2) - 1)
        int t = (acc + kFixOne) & (2 * kFixOne - 1);
        return fabsf((float)(t - kFixOne)) * kInv;
    }
    return fminf(fabsf((float)acc) * kInv, 1.0f);
}

// ---- accumulation of one clipped piece over the rows of a band --------------------------------------------
// (ax,ay)-(bx,by): piece after the reference's right-edge / x<0 handling.  Rows [row0,row1) of the canvas,
// columns [cx0, cx0+ncols) of it are this tile; `wc` is the reference's `width` (= img.width - 1).
__device__ __forceinline__ int to_fixed_f(float v) { return __float2int_rn(v * 16777216.0f); }

__device__ void accumulate_piece(double ax, double ay, double bx, double by, int row0, int row1, int cx0, int ncols, double wc,
                                 int* __restrict__ cells, int pitch, int* __restrict__ carry, int* __restrict__ touched) {
    if (fabs(ay - by) < kEps) return;  // src/rasterize.rs:400-403
    const int tile_end = cx0 + ncols;
    // x-extent of the piece decides how much work this tile has to do for it
    const double xmin = fmin(ax, bx), xmax = fmax(ax, bx);
    if (xmin >= (double)tile_end) return;  // entirely right of the tile: contributes nothing here
    float dirf = 1.0f;
    if (!(ay < by)) {  // src/rasterize.rs:405-409
        double t;
        t = ax; ax = bx; bx = t;
        t = ay; ay = by; by = t;
        dirf = -1.0f;
    }
    // rows of the reference loop (src/rasterize.rs:414, 421) intersected with the band
    const double ys = floor(fmax(ay, 0.0));
    const double ye = ceil(fmax(by, 0.0));
    const int rb = ys >= (double)row1 ? row1 : max(row0, (int)ys);
    const int re = ye >= (double)row1 ? row1 : (int)ye;
    if (rb >= re) return;
    if (xmax < (double)cx0 - 1.0) {
        // entirely left of the tile (with a pixel of slack for the span's last column): only the cover d = dir*dy
        // of every row reaches this tile, through the carry
        for (int y = rb; y < re; y++) {
            double dy = fmin((double)(y + 1), by) - fmax((double)y, ay);
            atomicAdd(&carry[y - row0], to_fixed_f(dirf * (float)dy));
        }
        return;
    }
    const double dxdy = (bx - ax) / (by - ay);
    const int wci = (int)wc;
    for (int y = rb; y < re; y++) {
        const double yt = fmax((double)y, ay);
        const double dy = fmin((double)(y + 1), by) - yt;
        const double x = ax + (yt - ay) * dxdy;  // the reference accumulates x row by row; this differs by rounding only
        const double xn = x + dxdy * dy;
        const double x0 = fmin(x, xn), x1 = fmax(x, xn);
        const double x0_floor = fmax(floor(x0), 0.0);
        const double x1_ceil = fmin(ceil(x1), wc);
        const int x0i = min(max((int)x0_floor, 0), wci);
        const int x1i = min(max((int)x1_ceil, 0), wci);
        if (x0i >= tile_end) continue;  // this row's span is right of the tile
        const int r = y - row0;
        const float d = dirf * (float)dy;
        const int fd = to_fixed_f(d);
        const bool narrow = x1i <= x0i + 1;
        const int last = narrow ? x0i + 1 : x1i;  // last column that receives a delta
        if (last < cx0) {                         // this row's span is left of the tile: only its cover arrives
            atomicAdd(&carry[r], fd);
            continue;
        }
        // Positions stay f64 (f32 ulp at x ~ 4096 would already exceed the 1e-4 budget); the fractional parts are
        // in [0,1] and the area polynomials are evaluated in f32 (error ~1e-7 of a pixel).
        float c0, sf = 0.f, a1 = 0.f, am = 0.f;
        const int n = x1i - x0i;
        if (narrow) {
            c0 = 1.0f - (float)(0.5 * (x + xn) - x0_floor);  // 1 - xmf, src/rasterize.rs:439
        } else {
            sf = 1.0f / (float)(x1 - x0);  // src/rasterize.rs:446-450
            const float x0f = (float)(x0 - x0_floor);
            const float x1f = (float)(x1 - x1_ceil + 1.0);
            c0 = 0.5f * sf * (1.0f - x0f) * (1.0f - x0f);
            am = 0.5f * sf * x1f * x1f;
            a1 = sf * (1.5f - x0f);
        }
        // coverage (as a fraction of d) of pixel x0i + j == running sum of the reference's deltas
        auto cov = [&](int j) -> float {
            if (j <= 0) return j == 0 ? c0 : 0.0f;
            if (narrow || j >= n) return 1.0f;
            if (j == n - 1) return 1.0f - am;
            return a1 + (float)(j - 1) * sf;
        };
        const int kb = max(x0i, cx0);
        const int ke = min(last, tile_end - 1);
        int prev = 0;
        if (kb > x0i) {  // the part of the span left of the tile collapses into the carry
            prev = to_fixed_f(d * cov(kb - 1 - x0i));
            atomicAdd(&carry[r], prev);
        }
        int* rowp = cells + r * pitch - cx0;
        for (int k = kb; k <= ke; k++) {
            const int cur = (k == last) ? fd : to_fixed_f(d * cov(k - x0i));
            const int diff = cur - prev;
            if (diff != 0) atomicAdd(&rowp[k], diff);
            prev = cur;
        }
        *touched = 1;
    }
}

// One flattened line -> the reference's clipping -> up to two pieces
__device__ void accumulate_line(const double4 l, int row0, int row1, int cx0, int ncols, double wc, int* cells, int pitch, int* carry,
                                int* touched) {
    double p0x = l.x, p0y = l.y, p1x = l.z, p1y = l.w;
    // src/rasterize.rs:370-387: lines crossing x == width
    if (p0x > wc || p1x > wc) {
        if (p0x > wc && p1x > wc) {
            p0x = wc - 0.001;
            p1x = wc - 0.001;
        } else {
            double t = (p0x - wc) / (p0x - p1x);
            double my = (1.0 - t) * p0y + t * p1y;
            if (p0x < wc) { p1x = wc; p1y = my; } else { p0x = wc; p0y = my; }
        }
    }
    // src/rasterize.rs:923-937 split_at_zero_x
    if (p0x >= 0.0 && p1x >= 0.0) {
        accumulate_piece(p0x, p0y, p1x, p1y, row0, row1, cx0, ncols, wc, cells, pitch, carry, touched);
    } else if (p0x <= 0.0 && p1x <= 0.0) {
        accumulate_piece(0.0, p0y, 0.0, p1y, row0, row1, cx0, ncols, wc, cells, pitch, carry, touched);
    } else {
        double t = p0x / (p0x - p1x);
        double mx = (1.0 - t) * p0x + t * p1x;
        double my = (1.0 - t) * p0y + t * p1y;
        if (p0x < 0.0) {
            accumulate_piece(mx, my, p1x, p1y, row0, row1, cx0, ncols, wc, cells, pitch, carry, touched);
            // rest = ((0, p0.y), mid) goes through the same function again in the reference
            if (mx <= 0.0) accumulate_piece(0.0, p0y, 0.0, my, row0, row1, cx0, ncols, wc, cells, pitch, carry, touched);
            else accumulate_piece(0.0, p0y, mx, my, row0, row1, cx0, ncols, wc, cells, pitch, carry, touched);
        } else {
            accumulate_piece(p0x, p0y, mx, my, row0, row1, cx0, ncols, wc, cells, pitch, carry, touched);
            if (mx <= 0.0) accumulate_piece(0.0, my, 0.0, p1y, row0, row1, cx0, ncols, wc, cells, pitch, carry, touched);
            else accumulate_piece(mx, my, 0.0, p1y, row0, row1, cx0, ncols, wc, cells, pitch, carry, touched);
        }
    }
}

template <int CW, int TH>
__global__ void __launch_bounds__(kThreads)
raster_kernel(const JobDev* __restrict__ jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first,
              const PaintDev* __restrict__ paints, const double4* __restrict__ lines, const uint32_t* __restrict__ band_offs,
              const uint32_t* __restrict__ refs, const Status* __restrict__ status) {
    constexpr int kPitch = CW + 4;
    __shared__ __align__(16) int cells[TH * kPitch];
    __shared__ int carry[TH];
    __shared__ int touched;
    __shared__ uint32_t s_job;
    __shared__ PaintDev s_paint;

    if (status->lines_overflow | status->refs_overflow | status->nan_flag | status->depth_flag) return;

    const int tid = threadIdx.x;
    uint32_t tile = tile_first + blockIdx.x;
    if (tid == 0) {
        s_job = job_first + find_job(n_jobs, tile, [&](uint32_t k) { return jobs[job_first + k].tile_begin; });
        touched = 0;
    }
    if (tid < TH) carry[tid] = 0;
    {
        int4 z = make_int4(0, 0, 0, 0);
        int4* c4 = reinterpret_cast<int4*>(cells);
        for (int i = tid; i < TH * kPitch / 4; i += kThreads) c4[i] = z;
    }
    __syncthreads();
    const JobDev& job = jobs[s_job];
    uint32_t lt = tile - job.tile_begin;
    int band = (int)(lt / job.n_chunks);
    int chunk = (int)(lt - (uint32_t)band * job.n_chunks);
    int row0 = band * TH;
    int row1 = min(row0 + TH, job.height);
    int cx0 = chunk * CW;
    double wc = job.clamp_w;
    int ncols = min(CW, (int)wc + 1 - cx0);  // columns that exist in the reference image (incl. the overflow column)
    const int mode = job.mode;
    const int rule = job.rule;

    if (mode == kModeFill && job.paint_index >= 0) {
        const int* src = reinterpret_cast<const int*>(&paints[job.paint_index]);
        int* dst = reinterpret_cast<int*>(&s_paint);
        for (int i = tid; i < (int)(sizeof(PaintDev) / 4); i += kThreads) dst[i] = src[i];
    }

    // ---- phase 1: accumulate the band's lines ----------------------------------------------------------
    uint32_t gb = job.band_begin + (uint32_t)band;
    uint32_t rbeg = band_offs[gb], rend = band_offs[gb + 1];
    if (ncols > 0) {
        for (uint32_t r = rbeg + tid; r < rend; r += kThreads) {
            double4 l = lines[refs[r]];
            accumulate_line(l, row0, row1, cx0, ncols, wc, cells, kPitch, carry, &touched);
        }
    }
    __syncthreads();

    // ---- phase 2: per-row scan, fill rule, store / composite --------------------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    const int wout = job.width_out;
    const bool any = touched != 0;
    const int nseg = min(CW / 128, (wout - cx0 + 127) >> 7);
    for (int r = warp; r < row1 - row0; r += kWarps) {
        int acc = carry[r];
        int* rowc = cells + r * kPitch;
        const int y = row0 + r;
        if (mode != kModeFill) {
            float* out = reinterpret_cast<float*>(job.canvas) + job.origin + (unsigned long long)y * job.row_stride + cx0;
            const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
            const int wrem = wout - cx0;  // columns of this row that exist from the tile's first column on
            for (int seg = 0; seg < nseg; seg++) {
                const int col = seg * 128 + lane * 4;
                int w0, w1, w2, w3;
                if (any) {
                    const int4 v = *reinterpret_cast<const int4*>(rowc + col);
                    const int p0 = v.x, p1 = p0 + v.y, p2 = p1 + v.z, p3 = p2 + v.w;
                    int incl = p3;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int nb = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += nb;
                    }
                    const int base = acc + incl - p3;
                    w0 = base + p0; w1 = base + p1; w2 = base + p2; w3 = base + p3;
                    acc += __shfl_sync(0xffffffffu, incl, 31);
                } else {
                    w0 = w1 = w2 = w3 = acc;
                }
                float4 cv = make_float4(coverage_from_fixed(w0, rule), coverage_from_fixed(w1, rule), coverage_from_fixed(w2, rule),
                                        coverage_from_fixed(w3, rule));
                if (mode == kModeCoverage) {  // mask_iter drops abs(alpha) < 1e-6 (src/rasterize.rs:348)
                    if (cv.x < 1e-6f) cv.x = 0.f;
                    if (cv.y < 1e-6f) cv.y = 0.f;
                    if (cv.z < 1e-6f) cv.z = 0.f;
                    if (cv.w < 1e-6f) cv.w = 0.f;
                }
                if (vec_ok && col + 3 < wrem) {
                    __stcs(reinterpret_cast<float4*>(out + col), cv);  // streaming 128-bit store, written once
                } else {
                    if (col < wrem) out[col] = cv.x;
                    if (col + 1 < wrem) out[col + 1] = cv.y;
                    if (col + 2 < wrem) out[col + 2] = cv.z;
                    if (col + 3 < wrem) out[col + 3] = cv.w;
                }
            }
        } else {
            float4* out = reinterpret_cast<float4*>(job.canvas) + job.origin + (unsigned long long)y * job.row_stride;
            for (int seg = 0; seg < nseg; seg++) {
                const int xs = cx0 + seg * 128;
                const int col = seg * 128 + lane * 4;
                int w0, w1, w2, w3;
                if (any) {
                    const int4 v = *reinterpret_cast<const int4*>(rowc + col);
                    const int p0 = v.x, p1 = p0 + v.y, p2 = p1 + v.z, p3 = p2 + v.w;
                    int incl = p3;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int nb = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += nb;
                    }
                    const int base = acc + incl - p3;
                    w0 = base + p0; w1 = base + p1; w2 = base + p2; w3 = base + p3;
                    acc += __shfl_sync(0xffffffffu, incl, 31);
                } else {
                    w0 = w1 = w2 = w3 = acc;
                }
                const float4 cv = make_float4(coverage_from_fixed(w0, rule), coverage_from_fixed(w1, rule),
                                              coverage_from_fixed(w2, rule), coverage_from_fixed(w3, rule));
                if (!__any_sync(0xffffffffu, fmaxf(fmaxf(cv.x, cv.y), fmaxf(cv.z, cv.w)) >= 1e-6f)) continue;  // nothing to paint
                // stage coverage so that consecutive lanes composite consecutive pixels (coalesced 16 B accesses)
                *reinterpret_cast<float4*>(rowc + col) = cv;
                __syncwarp();
                const float* covs = reinterpret_cast<const float*>(rowc + seg * 128);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int px = xs + i * 32 + lane;
                    const float alpha = covs[i * 32 + lane];
                    if (px < wout && alpha >= 1e-6f) {
                        float4 color = (job.paint_index >= 0) ? paint_at(s_paint, px, y) : make_float4(0.f, 0.f, 0.f, 0.f);
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
                __syncwarp();
            }
        }
    }
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

TileShape raster_tile_shape(int variant) {
    if (variant == 1) return TileShape{128, 64};
    return TileShape{1024, 8};
}

void launch_raster(int variant, const JobDev* jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first, uint32_t n_tiles,
                   const PaintDev* paints, const double4* lines, const uint32_t* band_offs, const uint32_t* refs,
                   const Status* status, cudaStream_t s) {
    if (n_tiles == 0) return;
    if (variant == 1)
        raster_kernel<128, 64><<<n_tiles, kThreads, 0, s>>>(jobs, n_jobs, job_first, tile_first, paints, lines, band_offs, refs, status);
    else
        raster_kernel<1024, 8><<<n_tiles, kThreads, 0, s>>>(jobs, n_jobs, job_first, tile_first, paints, lines, band_offs, refs, status);
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
