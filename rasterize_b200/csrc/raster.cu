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
#include "rgpu_internal.cuh"

namespace rgpu {

namespace {

constexpr double kEps = 2.220446049250313e-16;

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
template <bool EVENODD>
__device__ __forceinline__ float coverage_from_fixed(int acc) {
    constexpr float kInv = 1.0f / 16777216.0f;
    if (EVENODD) {
        // abs(((w + 1) rem_euclid 2) - 1)
        const int t = (acc + kFixOne) & (2 * kFixOne - 1);
        return fabsf((float)(t - kFixOne)) * kInv;
    }
    return fminf(fabsf((float)acc) * kInv, 1.0f);
}

// ---- shared-memory cell layout --------------------------------------------------------------------------
// In the scan phase lane l owns L = CW/32 consecutive columns of a row.  Cell (row r, tile column x) lives at
// r*pitch + swz<L>(x), swz<L>(x) = x + 4*(x/L): every run of L columns is followed by 4 padding ints, so the
// 128-bit accesses of 8 consecutive lanes (stride L+4 ints) fall on distinct bank quads — conflict-free both for
// the per-lane runs and (up to one 2-way pair for L = 16) for the transposed read-back used for coalesced stores.
template <int L>
__device__ __forceinline__ int swz(int x) { return x + ((x / L) << 2); }

__device__ __forceinline__ int to_fixed_f(float v) { return __float2int_rn(v * 16777216.0f); }

struct TileGeom {
    int row0, row1;   // canvas rows [row0, row1) of this band
    int cx0;          // first canvas column of the tile
    int tile_end;     // cx0 + number of reference columns in the tile (incl. the overflow column)
    int pitch;
    double wc;        // reference `width` (= img.width - 1)
    int wci;
};

// One (piece, row) span: the body of the reference's row loop (src/rasterize.rs:421-469) for canvas row y.
// (ax,ay) is the piece's upper end (ay < by), dirf = +-1.  Pixel coverages (running sums of the reference's
// deltas) are rounded to Q7.24 and the differences of consecutive rounded coverages are added to the cells of
// THIS tile only; their sum (a telescoping difference) goes to the row's tile total, from which the tiles to the
// right derive their carry-in.  Parts of the span in other tiles are added by those tiles (2-D bins).
template <int L>
__device__ __forceinline__ void span_row(double ax, double ay, double by, double dxdy, float dirf, int y, const TileGeom& g,
                                         int* __restrict__ cells, int* __restrict__ rowtot, int* __restrict__ row_touched) {
    const double yt = fmax((double)y, ay);
    const double dy = fmin((double)(y + 1), by) - yt;
    const double x = ax + (yt - ay) * dxdy;  // the reference accumulates x row by row; this differs by rounding only
    const double xn = x + dxdy * dy;
    const double x0 = fmin(x, xn), x1 = fmax(x, xn);
    const double x0_floor = fmax(floor(x0), 0.0);
    const double x1_ceil = fmin(ceil(x1), g.wc);
    const int x0i = min(max((int)x0_floor, 0), g.wci);
    const int x1i = min(max((int)x1_ceil, 0), g.wci);
    if (x0i >= g.tile_end) return;  // this row's span is right of the tile
    const int r = y - g.row0;
    const float d = dirf * (float)dy;
    const int fd = to_fixed_f(d);
    const bool narrow = x1i <= x0i + 1;
    const int last = narrow ? x0i + 1 : x1i;  // last column that receives a delta
    if (last < g.cx0) return;                 // this row's span is left of the tile: it arrives through the look-back
    // Positions stay f64 (f32 ulp at x ~ 4096 would already exceed the 1e-4 budget); the fractional parts are in
    // [0,1] and the area polynomials are evaluated in f32 (error ~1e-7 of a pixel).
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
    const int kb = max(x0i, g.cx0);
    const int ke = min(last, g.tile_end - 1);
    const int first = (kb > x0i) ? to_fixed_f(d * cov(kb - 1 - x0i)) : 0;  // rounded coverage just left of the tile
    int prev = first;
    int* rowp = cells + r * g.pitch;
    for (int k = kb; k <= ke; k++) {
        const int cur = (k == last) ? fd : to_fixed_f(d * cov(k - x0i));
        const int diff = cur - prev;
        if (diff != 0) atomicAdd(&rowp[swz<L>(k - g.cx0)], diff);
        prev = cur;
    }
    atomicAdd(&rowtot[r], prev - first);
    row_touched[r] = 1;
}

// Oriented piece ready for span_row, or nothing.  Returns the band rows [rb, re) it touches.
struct Piece {
    double ax, ay, by, dxdy;
    float dirf;
    int rb, re;
    int cls;  // 0 = nothing to do in this tile, 2 = needs span_row
};

__device__ __forceinline__ Piece classify_piece(double ax, double ay, double bx, double by, const TileGeom& g) {
    Piece p;
    p.cls = 0;
    p.rb = p.re = 0;
    p.dirf = 1.0f;
    p.dxdy = 0.0;
    if (fabs(ay - by) < kEps) { p.ax = ax; p.ay = ay; p.by = by; return p; }  // src/rasterize.rs:400-403
    const double xmin = fmin(ax, bx), xmax = fmax(ax, bx);
    if (!(ay < by)) {  // src/rasterize.rs:405-409
        double t;
        t = ax; ax = bx; bx = t;
        t = ay; ay = by; by = t;
        p.dirf = -1.0f;
    }
    p.ax = ax; p.ay = ay; p.by = by;
    if (xmin >= (double)g.tile_end) return p;    // entirely right of the tile: contributes nothing here
    if (xmax < (double)g.cx0 - 1.0) return p;    // entirely left (a pixel of slack for the span's last column): look-back
    // rows of the reference loop (src/rasterize.rs:414, 421) intersected with the band
    const double ys = floor(fmax(ay, 0.0));
    const double ye = ceil(fmax(by, 0.0));
    p.rb = ys >= (double)g.row1 ? g.row1 : max(g.row0, (int)ys);
    p.re = ye >= (double)g.row1 ? g.row1 : (int)ye;
    if (p.rb >= p.re) return p;
    p.dxdy = (bx - ax) / (by - ay);
    p.cls = 2;
    return p;
}

template <int L>
__device__ void piece_serial(double ax, double ay, double bx, double by, const TileGeom& g, int* cells, int* rowtot, int* row_touched) {
    const Piece p = classify_piece(ax, ay, bx, by, g);
    if (p.cls == 2)
        for (int y = p.rb; y < p.re; y++) span_row<L>(p.ax, p.ay, p.by, p.dxdy, p.dirf, y, g, cells, rowtot, row_touched);
}

// ---- carry look-back state: one 64-bit word per (tile, row): [63:34] epoch, [33:32] flag, [31:0] value ---------
constexpr unsigned long long kFlagAgg = 1ull << 32;     // value = this tile's row total
constexpr unsigned long long kFlagPrefix = 2ull << 32;  // value = inclusive prefix over the tiles of the band so far
__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

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
            const float c = coverage_from_fixed<EVENODD>(acc);  // no line touched this row of the tile: constant coverage
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
              const PaintDev* __restrict__ paints, const uint32_t* __restrict__ tile_offs, const double4* __restrict__ bin_lines,
              unsigned long long* __restrict__ tile_state, uint32_t epoch, uint32_t* __restrict__ ticket,
              const Status* __restrict__ status) {
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
    {
        const int4 z = make_int4(0, 0, 0, 0);
        int4* c4 = reinterpret_cast<int4*>(cells);
        for (int i = tid; i < TH * Cfg::kPitch / 4; i += THREADS) c4[i] = z;
    }
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
    // 1a: one line per lane — the reference's clipping, orientation and row range; the (piece,row) spans of the
    //     32 lines are compacted into the warp's span list (warp prefix sum, no atomics)
    // 1b: one lane per span — no row-loop divergence
    unsigned short* spans = spans_all + warp * Cfg::kWarpSpanCap;
    const uint32_t rbeg = tile_offs[tile], rend = tile_offs[tile + 1];
    for (uint32_t r0 = rbeg + warp * 32; r0 < rend; r0 += THREADS) {
        const uint32_t r = r0 + lane;
        Piece p;
        p.cls = 0;
        p.rb = p.re = 0;
        if (r < rend) {
            const double4 l = bin_lines[r];
            double p0x = l.x, p0y = l.y, p1x = l.z, p1y = l.w;
            // src/rasterize.rs:370-387: lines crossing x == width
            if (p0x > wc || p1x > wc) {
                if (p0x > wc && p1x > wc) {
                    p0x = wc - 0.001;
                    p1x = wc - 0.001;
                } else {
                    const double t = (p0x - wc) / (p0x - p1x);
                    const double my = (1.0 - t) * p0y + t * p1y;
                    if (p0x < wc) { p1x = wc; p1y = my; } else { p0x = wc; p0y = my; }
                }
            }
            // src/rasterize.rs:923-937 split_at_zero_x
            if (p0x < 0.0 || p1x < 0.0) {
                if (p0x <= 0.0 && p1x <= 0.0) {
                    p0x = 0.0;
                    p1x = 0.0;
                } else {
                    const double t = p0x / (p0x - p1x);
                    const double mx = (1.0 - t) * p0x + t * p1x;
                    const double my = (1.0 - t) * p0y + t * p1y;
                    // the outside part, folded onto x = 0, goes through the same function again in the reference;
                    // rare (only lines crossing the left edge): done serially by this lane
                    if (p0x < 0.0) {
                        if (mx <= 0.0) piece_serial<L>(0.0, p0y, 0.0, my, g, cells, rowtot, row_touched);
                        else piece_serial<L>(0.0, p0y, mx, my, g, cells, rowtot, row_touched);
                        p0x = mx; p0y = my;
                    } else {
                        if (mx <= 0.0) piece_serial<L>(0.0, my, 0.0, p1y, g, cells, rowtot, row_touched);
                        else piece_serial<L>(mx, my, 0.0, p1y, g, cells, rowtot, row_touched);
                        p1x = mx; p1y = my;
                    }
                }
            }
            p = classify_piece(p0x, p0y, p1x, p1y, g);
        }
        int n = (p.cls == 2) ? p.re - p.rb : 0;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int base = incl - n;
        if (n > 0) {
            if (base + n <= Cfg::kWarpSpanCap) {
                p_ax[tid] = p.ax; p_ay[tid] = p.ay; p_by[tid] = p.by; p_dxdy[tid] = p.dxdy; p_dir[tid] = p.dirf;
                for (int k = 0; k < n; k++) spans[base + k] = (unsigned short)((lane << Cfg::kRowBits) | (p.rb + k - row0));
            } else {  // list full (only possible for TH = 64): do the rows here
                for (int y = p.rb; y < p.re; y++) span_row<L>(p.ax, p.ay, p.by, p.dxdy, p.dirf, y, g, cells, rowtot, row_touched);
            }
        }
        __syncwarp();
        // lanes are in list order: everything before the first lane that did not fit is in the list
        const unsigned nofit = __ballot_sync(0xffffffffu, n > 0 && base + n > Cfg::kWarpSpanCap);
        const int ns = nofit ? __shfl_sync(0xffffffffu, base, __ffs(nofit) - 1) : total;
        const int wbase = warp * 32;
        for (int i = lane; i < ns; i += 32) {
            const int e = spans[i];
            const int slot = wbase + (e >> Cfg::kRowBits);
            const int y = row0 + (e & ((1 << Cfg::kRowBits) - 1));
            span_row<L>(p_ax[slot], p_ay[slot], p_by[slot], p_dxdy[slot], p_dir[slot], y, g, cells, rowtot, row_touched);
        }
        __syncwarp();
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
                            uint32_t n_tiles, const PaintDev* paints, const uint32_t* tile_offs, const double4* bin_lines,
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
                                                                  bin_lines, tile_state, epoch, ticket, status);
}

TileShape raster_tile_shape(int variant) {
    if (variant == 1) return TileShape{128, 64};
    return TileShape{512, 8};
}

void launch_raster(int variant, const JobDev* jobs, const JobDev* h_jobs, uint32_t n_jobs, uint32_t job_first, uint32_t tile_first,
                   uint32_t n_tiles, const PaintDev* paints, const uint32_t* tile_offs, const double4* bin_lines,
                   unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, const Status* status, cudaStream_t s) {
    if (n_tiles == 0) return;
    if (variant == 1)
        launch_raster_t<128, 64, 256>(jobs, h_jobs, n_jobs, job_first, tile_first, n_tiles, paints, tile_offs, bin_lines, tile_state, epoch,
                                      ticket, status, s);
    else
        launch_raster_t<512, 8, 128>(jobs, h_jobs, n_jobs, job_first, tile_first, n_tiles, paints, tile_offs, bin_lines, tile_state, epoch,
                                     ticket, status, s);
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
