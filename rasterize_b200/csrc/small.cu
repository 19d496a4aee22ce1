// Fused fill path for small canvases (at most 64 x 64 visible pixels, e.g. glyph batches — BASELINE config 4):
// ONE 160-thread CTA per job runs K1..K4 end to end.  No line buffer, no bins and no carries ever touch HBM: traffic
// is the algorithmic minimum (control points in, pixels out), and a whole batch is one launch.
//
// Round-2 design (round 1: one 256-thread CTA, a depth-first walk with a 1.5 KB local-memory stack per slot thread, block
// barriers between flattening, accumulation and the scan: 1.40 ms per 20 000 glyphs against 0.94 ms now).  The kernel is bound by
// instruction issue, not by HBM, so everything below is about executing few warp-instructions with full warps:
//   * curves are transformed ONCE into shared memory (chunks of 20 curves); a curve's subdivision tree is cut at
//     depth 3 (8 slot threads per curve, 20 curves = 160 threads).  A slot thread descends to its subtree root and
//     walks the two levels below it with straight-line code in registers — no stack, no loop, every lane of the warp
//     on the same instruction — testing each child's flatness as soon as it exists;
//   * leaves are parked in a per-warp shared-memory line queue through warp-collective "emit sites" (ballot +
//     popc), nodes that are still not flat two levels below the slot root (a tenth of a glyph's level-5 nodes) go to
//     a small per-warp node queue and are expanded later with one lane per child; anything deeper takes a stack-free
//     walk by path bits (the next node of the depth-first order is recomputed from its subtree root);
//   * a warp's line queue is drained once per round (about 120 lines of a glyph), lane = line: orientation, slope and
//     row range, then the line's rows in a loop whose trip count is the warp's longest line (a glyph's lines cover 2.0
//     rows on average, at most 4 for 99 %).  A span over at most two columns (93 % of a glyph's spans) is three
//     shared-memory integer reductions computed without a branch; the few wider ones are set aside and done together
//     at the end of the pass, so the rare path does not run in every iteration with two lanes.  (A span list built by
//     a warp prefix sum and processed with lane = span keeps more lanes busy but costs as many instructions again for
//     its bookkeeping: 0.96 vs 0.87 ms per 20 000 glyphs.);
//   * there are two block barriers per job (after staging, before the row scan);
//   * end points of a small canvas are below 128, where f32 resolves 7.6e-6: a queued line's spans are evaluated in
//     f32 (lines that leave the canvas columns or lie far outside its rows take the f64 path of the tiled kernels,
//     one lane each — rare).  Consecutive lines share their (identically rounded) end points, so contours stay
//     closed and the winding stays exact; the coverage error is bounded by the 4e-6 shift of an end point.
//
// Same arithmetic as the tiled pipeline otherwise: subdivision decisions are bit-exact f64 in the reference's
// expression order (products by powers of two are exact, so `fma(0.5, a, b)` is `0.5 * a + b` with one rounding —
// only the inexact products are rounded separately), Q7.24 telescoping coverage differences, integer atomics (the
// result does not depend on the order in which lines are accumulated).
// Reference citations: flatten src/path.rs:744-795, src/curve.rs:413-417, 449-456, 692-697, 731-747; raster
// src/rasterize.rs:365-507; fill src/rasterize.rs:70-115.
#include "flatten_device.cuh"
#include "raster_device.cuh"

#include <algorithm>
#include <type_traits>

namespace rgpu {

namespace {

using namespace fl;
using namespace rs;

constexpr int kGThreads = 160;
constexpr int kGWarps = kGThreads / 32;
constexpr int kGDepth = 3;                                      // slot cut: 8 slot threads per curve
constexpr int kGSlots = 1 << kGDepth;
constexpr int kGCurves = kGThreads / kGSlots;                   // curves staged in shared memory at a time
#ifndef RGPU_GQUEUE
#define RGPU_GQUEUE 160
#endif
constexpr int kGQueue = RGPU_GQUEUE;                                    // per-warp line queue: a round of 32 slots leaves ~120 lines of a glyph
constexpr int kGIds = 8;                                        // per-warp scratch words (the count of deferred wide spans)
#ifndef RGPU_TWO_CLASS
#define RGPU_TWO_CLASS 1
#endif
#ifndef RGPU_GTALL
#define RGPU_GTALL 6
#endif
constexpr int kGTall = RGPU_GTALL;                              // a line over more rows than this is spread over the lanes (deferred list)
constexpr int kGWide = 64;                                      // spans over three or more columns, deferred to the end of a pass
constexpr int kGDeep = 16;                                      // per-warp queue of nodes not flat at level kGDepth + 2
constexpr int kGMaxBelow = kSlotDepth + kMaxStack - kGDepth - 3;  // walk_deep starts three levels below a slot root
constexpr int kSmMaxW = 64, kSmMaxH = 64;
constexpr int kSmPitch = 68;                                    // 64 columns + overflow column, padded to a multiple of 4 ints
constexpr unsigned kFull = 0xffffffffu;

// A curve node in registers (quads use points 0..2)
struct Nd {
    double x0, x1, x2, x3, y0, y1, y2, y3;
};

__device__ __forceinline__ bool nd_has_nan(const Nd& n, int kind) {
    bool b = isnan(n.x0) || isnan(n.y0) || isnan(n.x1) || isnan(n.y1) || isnan(n.x2) || isnan(n.y2);
    if (kind == 4) b = b || isnan(n.x3) || isnan(n.y3);
    return b;
}

// Curve::flatness, src/curve.rs:413-417 (quad), 692-697 (cubic).  2 * p is exact, so `3 p1 - 2 p0` is one fma.
__device__ __forceinline__ double nd_flatness(const Nd& n, int kind) {
    if (kind == 4) {
        const double ux = dsub(fma(-2.0, n.x0, dmul(3.0, n.x1)), n.x3);
        const double uy = dsub(fma(-2.0, n.y0, dmul(3.0, n.y1)), n.y3);
        const double vx = fma(-2.0, n.x3, dsub(dmul(3.0, n.x2), n.x0));
        const double vy = fma(-2.0, n.y3, dsub(dmul(3.0, n.y2), n.y0));
        return dadd(dmax(dmul(ux, ux), dmul(vx, vx)), dmax(dmul(uy, uy), dmul(vy, vy)));
    }
    // 2.0 * p1 - p0 - p2
    const double dx = dsub(fma(2.0, n.x1, -n.x0), n.x2);
    const double dy = dsub(fma(2.0, n.y1, -n.y0), n.y2);
    return dadd(dmul(dx, dx), dmul(dy, dy));
}

// split() of one coordinate (src/curve.rs:449-456, 731-747): both halves.
//   mid = 0.125 p0 + 0.375 p1 + 0.375 p2 + 0.125 p3 (left to right)
//   left  = (p0, 0.5 p0 + 0.5 p1, 0.25 p0 + 0.5 p1 + 0.25 p2, mid)
//   right = (mid, 0.25 p1 + 0.5 p2 + 0.25 p3, 0.5 p2 + 0.5 p3, p3)
__device__ __forceinline__ void split_cubic(double p0, double p1, double p2, double p3, double& a1, double& a2, double& mid, double& b1,
                                            double& b2) {
    mid = fma(0.125, p3, dadd(fma(0.125, p0, dmul(0.375, p1)), dmul(0.375, p2)));
    const double h1 = dmul(0.5, p1), h2 = dmul(0.5, p2);
    a1 = fma(0.5, p0, h1);
    a2 = fma(0.25, p2, fma(0.25, p0, h1));
    b1 = fma(0.25, p3, fma(0.25, p1, h2));
    b2 = fma(0.5, p3, h2);
}
// quad: mid = 0.25 * (p0 + 2 p1 + p2); left = (p0, 0.5 (p0 + p1), mid); right = (mid, 0.5 (p1 + p2), p2)
__device__ __forceinline__ void split_quad(double p0, double p1, double p2, double& a1, double& mid, double& b1) {
    mid = dmul(0.25, dadd(fma(2.0, p1, p0), p2));
    a1 = dmul(0.5, dadd(p0, p1));
    b1 = dmul(0.5, dadd(p1, p2));
}
__device__ __forceinline__ void nd_split(const Nd& n, int kind, Nd& a, Nd& b) {
    if (kind == 4) {
        double mx, my;
        split_cubic(n.x0, n.x1, n.x2, n.x3, a.x1, a.x2, mx, b.x1, b.x2);
        split_cubic(n.y0, n.y1, n.y2, n.y3, a.y1, a.y2, my, b.y1, b.y2);
        a.x0 = n.x0; a.y0 = n.y0; a.x3 = mx; a.y3 = my;
        b.x0 = mx; b.y0 = my; b.x3 = n.x3; b.y3 = n.y3;
    } else {
        double mx, my;
        split_quad(n.x0, n.x1, n.x2, a.x1, mx, b.x1);
        split_quad(n.y0, n.y1, n.y2, a.y1, my, b.y1);
        a.x0 = n.x0; a.y0 = n.y0; a.x2 = mx; a.y2 = my;
        b.x0 = mx; b.y0 = my; b.x2 = n.x2; b.y2 = n.y2;
        a.x3 = a.y3 = b.x3 = b.y3 = 0.0;
    }
}
__device__ __forceinline__ void nd_child(Nd& n, int kind, bool r) {
    Nd a, b;
    nd_split(n, kind, a, b);
    n.x0 = r ? b.x0 : a.x0; n.x1 = r ? b.x1 : a.x1; n.x2 = r ? b.x2 : a.x2; n.x3 = r ? b.x3 : a.x3;
    n.y0 = r ? b.y0 : a.y0; n.y1 = r ? b.y1 : a.y1; n.y2 = r ? b.y2 : a.y2; n.y3 = r ? b.y3 : a.y3;
}
// One half of split() with the side known at compile time (the straight-line walk): 10 f64 operations per coordinate.
template <bool R>
__device__ __forceinline__ void half_cubic(double p0, double p1, double p2, double p3, double& o0, double& o1, double& o2, double& o3) {
    const double mid = fma(0.125, p3, dadd(fma(0.125, p0, dmul(0.375, p1)), dmul(0.375, p2)));
    if (R) {
        const double h2 = dmul(0.5, p2);
        o0 = mid; o1 = fma(0.25, p3, fma(0.25, p1, h2)); o2 = fma(0.5, p3, h2); o3 = p3;
    } else {
        const double h1 = dmul(0.5, p1);
        o0 = p0; o1 = fma(0.5, p0, h1); o2 = fma(0.25, p2, fma(0.25, p0, h1)); o3 = mid;
    }
}
template <bool R>
__device__ __forceinline__ Nd nd_half(const Nd& n, int kind) {
    Nd o;
    if (kind == 4) {
        half_cubic<R>(n.x0, n.x1, n.x2, n.x3, o.x0, o.x1, o.x2, o.x3);
        half_cubic<R>(n.y0, n.y1, n.y2, n.y3, o.y0, o.y1, o.y2, o.y3);
    } else {
        double ax, mx, bx, ay, my, by;
        split_quad(n.x0, n.x1, n.x2, ax, mx, bx);
        split_quad(n.y0, n.y1, n.y2, ay, my, by);
        o.x0 = R ? mx : n.x0; o.x1 = R ? bx : ax; o.x2 = R ? n.x2 : mx; o.x3 = 0.0;
        o.y0 = R ? my : n.y0; o.y1 = R ? by : ay; o.y2 = R ? n.y2 : my; o.y3 = 0.0;
    }
    return o;
}
__device__ __forceinline__ double nd_endx(const Nd& n, int kind) { return kind == 4 ? n.x3 : n.x2; }
__device__ __forceinline__ double nd_endy(const Nd& n, int kind) { return kind == 4 ? n.y3 : n.y2; }

__device__ __forceinline__ int fix_f(float v, float scale) { return __float2int_rn(v * scale); }
// 1 / x for x in (1e-20, 1e3): the bare MUFU.RCP (1 ulp), without the range fix-up of __fdividef
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- shared memory of a CTA: one dynamic block, addressed through these views so that the out-of-line helpers see
// shared-memory addresses ---------------------------------------------------------------------------------
constexpr int kCells = kSmMaxH * kSmPitch;
constexpr size_t kOffCells = 0;
constexpr size_t kOffLineq = kOffCells + sizeof(int) * kCells;          // per-warp line queues (phase 2 of the one-glyph kernel: a gradient's table)
constexpr size_t kOffDq = kOffLineq + sizeof(float4) * kGWarps * kGQueue;   // per-warp deep-node queues
constexpr size_t kOffRoots = kOffDq + sizeof(double) * kGWarps * kGDeep * 8;  // transformed control points: [20][4] x, then y
constexpr size_t kOffIds = kOffRoots + sizeof(double) * kGCurves * 4 * 2;
constexpr size_t kOffWide = kOffIds + sizeof(unsigned short) * kGWarps * kGIds;
constexpr size_t kOffWideY = kOffWide + sizeof(float4) * kGWarps * kGWide;
constexpr size_t kOffDkind = kOffWideY + kGWarps * kGWide;
constexpr size_t kOffMeta = kOffDkind + 16 * ((kGWarps * kGDeep + 15) / 16);
constexpr size_t kOffRowtot = kOffMeta + 48;
constexpr size_t kOffJob = kOffRowtot + sizeof(int) * 2 * kSmMaxH;
constexpr size_t kJobPitch = (sizeof(JobDev) + 15) / 16 * 16;
constexpr size_t kOffBars = kOffJob + kJobPitch;
constexpr size_t kSmemBytes = kOffBars + 64;
static_assert(sizeof(PaintDev) <= sizeof(float4) * kGWarps * kGQueue, "paint overlay");
extern __shared__ __align__(16) unsigned char sm_raw[];
#define s_cells (reinterpret_cast<int*>(sm_raw + kOffCells))
#define s_lineq (reinterpret_cast<float4(*)[kGQueue]>(sm_raw + kOffLineq))
#define s_dq (reinterpret_cast<double(*)[kGDeep * 8]>(sm_raw + kOffDq))
#define s_x (reinterpret_cast<double(*)[4]>(sm_raw + kOffRoots))
#define s_y (reinterpret_cast<double(*)[4]>(sm_raw + kOffRoots + sizeof(double) * kGCurves * 4))
#define s_ids (reinterpret_cast<unsigned short(*)[kGIds]>(sm_raw + kOffIds))
#define s_wide_p (reinterpret_cast<float4(*)[kGWide]>(sm_raw + kOffWide))
#define s_wide_y (reinterpret_cast<unsigned char(*)[kGWide]>(sm_raw + kOffWideY))
#define s_dkind (reinterpret_cast<unsigned char(*)[kGDeep]>(sm_raw + kOffDkind))
#define s_meta (reinterpret_cast<unsigned char*>(sm_raw + kOffMeta))
#define s_rowtot (reinterpret_cast<int*>(sm_raw + kOffRowtot))
#define s_row_touched (reinterpret_cast<int*>(sm_raw + kOffRowtot) + kSmMaxH)
#define s_job (*reinterpret_cast<JobDev*>(sm_raw + kOffJob))
#define s_nlines (*reinterpret_cast<uint32_t*>(sm_raw + kOffBars + 32))

// Per-job constants of the accumulation
struct Canvas {
    int H, wci, tile_end;
    float wcf;
    float fix_scale;  // 2^fix_shift of the batch (raster_device.cuh: Fix)
    double wc;
};

__device__ __forceinline__ void cell_add(int idx, int v) {
    // result unused: a shared-memory reduction
    atomicAdd(&s_cells[idx], v);
}

// A (line, row) span of a prepared line p = (ax, ay, +-by, dxdy) (ay < by, the sign of the third field is the line's
// direction): the body of the reference's row loop (src/rasterize.rs:421-469) in f32.
struct Span {
    float xt, xn, xa, xb, d;
    int x0i, x1i, n, fd, at;
};
__device__ __forceinline__ Span span_head(const float4 p, const int y, const Canvas& cv) {
    Span s;
    const float ax = p.x, ay = p.y, by = fabsf(p.z), dxdy = p.w;
    const float fy = (float)y;
    const float yt = fmaxf(fy, ay), yb = fminf(fy + 1.0f, by);
    s.d = copysignf(yb - yt, p.z);
    s.xt = fminf(fmaxf(fmaf(yt - ay, dxdy, ax), 0.0f), cv.wcf);
    s.xn = fminf(fmaxf(fmaf(yb - ay, dxdy, ax), 0.0f), cv.wcf);
    s.xa = fminf(s.xt, s.xn);
    s.xb = fmaxf(s.xt, s.xn);
    s.x0i = (int)s.xa;
    s.x1i = min((int)ceilf(s.xb), cv.wci);
    s.n = s.x1i - s.x0i;
    s.fd = fix_f(s.d, cv.fix_scale);
    s.at = y * kSmPitch + s.x0i;
    return s;
}
// one column (src/rasterize.rs:437-444) or two (:445-459), without a branch between them
__device__ __forceinline__ void span_narrow(const Span& s, const Canvas& cv) {
    const float fx0 = (float)s.x0i;
    const float sf = rcp_fast(fmaxf(s.xb - s.xa, 1e-20f));
    const float x1f = s.xb - (float)s.x1i + 1.0f;
    const float c_narrow = 1.0f - (0.5f * (s.xt + s.xn) - fx0);
    const float u = 1.0f - (s.xa - fx0);
    const float c_wide = 0.5f * sf * u * u;
    const bool two = s.n == 2;
    const int qa = fix_f(s.d * (two ? c_wide : c_narrow), cv.fix_scale);
    const int qb = two ? fix_f(s.d * (1.0f - 0.5f * sf * x1f * x1f), cv.fix_scale) : s.fd;
    cell_add(s.at, qa);
    if (s.x0i + 1 < cv.tile_end) cell_add(s.at + 1, qb - qa);
    if (two) cell_add(s.at + 2, s.fd - qb);  // x0i + 2 == x1i <= wci < tile_end
}
// three or more columns (src/rasterize.rs:445-468)
__device__ __forceinline__ void span_wide(const Span& s, const Canvas& cv) {
    const float x0f = s.xa - (float)s.x0i;
    const float sf = rcp_fast(s.xb - s.xa);
    const float x1f = s.xb - (float)s.x1i + 1.0f;
    const float c0 = 0.5f * sf * (1.0f - x0f) * (1.0f - x0f);
    const float cl = 1.0f - 0.5f * sf * x1f * x1f;
    const float a1 = sf * (1.5f - x0f);
    int prev = fix_f(s.d * c0, cv.fix_scale);
    cell_add(s.at, prev);
    for (int j = 1; j < s.n; j++) {
        const float c = (j == s.n - 1) ? cl : a1 + (float)(j - 1) * sf;
        const int cur = fix_f(s.d * c, cv.fix_scale);
        cell_add(s.at + j, cur - prev);
        prev = cur;
    }
    cell_add(s.at + s.n, s.fd - prev);  // x1i <= wci < tile_end
}

// Accumulate this warp's queue: `counts` = short lines (at the front) | other lines (at the back) << 16; a group of 32 never
// mixes the two.  Every end point lies inside the canvas columns [0, wc] and
// within (-64, 128) of its rows, so the clipping branches of signed_difference_line (x > width, x < 0) cannot trigger.
// Lane = line, 32 at a time: orientation, slope and row range (src/rasterize.rs:400-421), then the rows of the line in a loop
// whose trip count is the warp's longest line (a glyph's lines cover 2.0 rows on average, 4 or fewer for 99 %: the loop runs
// about four times per round with predicated lanes, and there is no span list to build and read back).  Spans over three or
// more columns (7 %) are set aside with their prepared line and done together at the end of the pass, and so are all rows of a
// line taller than kGTall rows (each glyph has one: its closing line), which would otherwise set the trip count for 32 lanes.
// Measured and not kept: lines over one or two rows done in two predicated steps and the longer ones compacted for a second
// pass (more code, same instruction count: 0.93 vs 0.87 ms per 20 000 glyphs).
__device__ __noinline__ void accumulate_warp(const int counts, const Canvas cv) {
    const int count_a = counts & 0xffff, count_b = counts >> 16;  // short lines from the front, the others from the back
    const int groups_a = (count_a + 31) >> 5, groups = groups_a + ((count_b + 31) >> 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* __restrict__ q = s_lineq[warp];
    float4* __restrict__ wide_p = s_wide_p[warp];
    unsigned char* __restrict__ wide_y = s_wide_y[warp];
    int* const n_wide = reinterpret_cast<int*>(s_ids[warp]);  // count of deferred wide spans
    if (lane == 0) *n_wide = 0;
    __syncwarp();
    // the deferred list: (prepared line, row) spans — those over three or more columns, and every row of a tall line.
    // (Measured and not kept: spans over more than 8 / 16 columns done by the whole warp with lane = column — no change.)
    auto do_wide = [&]() {
        __syncwarp();
        const int nw = min(*n_wide, kGWide);
        __syncwarp();
        for (int k = lane; k < nw; k += 32) {
            const Span s = span_head(wide_p[k], wide_y[k], cv);
            if (s.n <= 2) span_narrow(s, cv);
            else span_wide(s, cv);
        }
        if (lane == 0) *n_wide = 0;
        __syncwarp();
    };
    for (int g = 0; g < groups; g++) {
        const bool back = g >= groups_a;
        const int i = ((back ? g - groups_a : g) << 5) + lane;
        int n = 0, rb = 0;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < (back ? count_b : count_a)) {
            const float4 l = q[back ? kGQueue - 1 - i : i];
            float ax = l.x, ay = l.y, bx = l.z, by = l.w;
            float dirf = 1.0f;
            if (ay > by) {
                float t;
                t = ax; ax = bx; bx = t;
                t = ay; ay = by; by = t;
                dirf = -1.0f;
            }
            if (ay != by) {
                p = make_float4(ax, ay, copysignf(by, dirf), __fdividef(bx - ax, by - ay));
                rb = (int)fmaxf(ay, 0.0f);
                n = max(min(cv.H, (int)ceilf(by)) - rb, 0);  // lines above or below the canvas: no rows
            }
        }
        // A tall line (a glyph has one or two: the closing line, a long straight side) would hold the whole warp in the row
        // loop with one active lane: its rows go to the deferred list instead, 32 lanes writing 32 rows at a time.
        unsigned tall = __ballot_sync(kFull, n > kGTall);
        const int n_main = n > kGTall ? 0 : n;
        const int n_max = __reduce_max_sync(kFull, n_main);
        auto row = [&](const int k) {
            if (k < n_main) {
                const Span s = span_head(p, rb + k, cv);
                if (s.n <= 2) {
                    span_narrow(s, cv);
                } else {
                    const int slot = atomicAdd(n_wide, 1);
                    if (slot < kGWide) {
                        wide_p[slot] = p;
                        wide_y[slot] = (unsigned char)(rb + k);
                    } else {
                        span_wide(s, cv);  // list full: here and now
                    }
                }
            }
        };
        for (int k = 0; k < n_max; k++) row(k);  // (unrolled by two: 3.69 vs 3.63 ms per 100 000 glyphs — not kept)
        __syncwarp();
        while (tall) {
            const int src = __ffs(tall) - 1;
            tall &= tall - 1;
            const float4 pl = make_float4(__shfl_sync(kFull, p.x, src), __shfl_sync(kFull, p.y, src), __shfl_sync(kFull, p.z, src),
                                          __shfl_sync(kFull, p.w, src));
            const int rbl = __shfl_sync(kFull, rb, src), nl = __shfl_sync(kFull, n, src);  // nl <= 64 = kGWide
            if (*n_wide + nl > kGWide) do_wide();
            const int base = min(*n_wide, kGWide);
            __syncwarp();
            for (int k = lane; k < nl; k += 32) {
                wide_p[base + k] = pl;
                wide_y[base + k] = (unsigned char)(rbl + k);
            }
            if (lane == 0) *n_wide = base + nl;
            __syncwarp();
        }
        if (*n_wide > kGWide - 32) do_wide();  // warp-uniform: every lane reads the same word after the loop
    }
    do_wide();
}

// rare paths, kept out of line so that their registers do not count against the flatten walk
__device__ __noinline__ void line_slow(double x0, double y0, double x1, double y1, const Canvas cv) {
    if (fmax(y0, y1) <= 0.0 || fmin(y0, y1) >= (double)cv.H) return;  // misses every row: the reference's row loop is empty
    TileGeom g;
    g.row0 = 0;
    g.row1 = cv.H;
    g.cx0 = 0;
    g.wc = cv.wc;
    g.wci = cv.wci;
    g.tile_end = cv.tile_end;
    g.pitch = kSmPitch;
    g.fix_scale = cv.fix_scale;
    line_serial<false>(make_double4(x0, y0, x1, y1), g, s_cells, s_rowtot, s_row_touched);
}

// All coordinates inside the canvas columns and near its rows: the f32 spans apply (see accumulate_warp).  A curve whose
// control points pass this test has all its leaves pass it (they lie in the convex hull).
__device__ __forceinline__ bool point_safe(double x, double y, const Canvas& cv) {
    return x >= 0.0 && x <= cv.wc && y > -64.0 && y < 128.0;
}

// Per-warp state of the flatten walk
struct Warp {
    int qn, dn;            // warp-uniform fill levels of the line queue / the deep-node queue.  qn = short | tall << 16: lines over
                           // at most two rows fill the queue from the front, the others from the back (see emit_site)
    uint32_t lines;        // leaves found by this lane
    unsigned lt_mask;
    int warp;
};

// Warp-collective emit site: lanes with `pred` hold a leaf (x0, y0) -> (x1, y1) of a "safe" curve.  No call, no drain:
// the callers drain at points where little is live and size the queue for what can arrive in between.
__device__ __forceinline__ void emit_site(bool pred, double x0, double y0, double x1, double y1, Warp& w) {
    const float fx0 = (float)x0, fy0 = (float)y0, fx1 = (float)x1, fy1 = (float)y1;
#if RGPU_TWO_CLASS
    // Two classes, kept apart in the queue: the row loop of accumulate_warp runs to the longest line of a group of 32, and
    // 75 % of a glyph's lines cover one or two rows — mixed with the others they idle through the long ones' iterations
    // (measured: 12.9 of 32 lanes active).  Sorted into two classes a warp's ~110 lines take 57 instead of 81 loop iterations
    // per glyph (as many as a full sort by row count would give).  The test is only a hint: any assignment is correct.
    const bool big = ceilf(fmaxf(fy0, fy1)) - floorf(fminf(fy0, fy1)) > 2.0f;
    const unsigned ma = __ballot_sync(kFull, pred && !big), mb = __ballot_sync(kFull, pred && big);
    if (pred) {
        w.lines++;
        const int pos = big ? kGQueue - 1 - ((w.qn >> 16) + __popc(mb & w.lt_mask)) : (w.qn & 0xffff) + __popc(ma & w.lt_mask);
        s_lineq[w.warp][pos] = make_float4(fx0, fy0, fx1, fy1);
    }
    w.qn += __popc(ma) + (__popc(mb) << 16);
#else
    const unsigned mf = __ballot_sync(kFull, pred);
    if (pred) {
        w.lines++;
        s_lineq[w.warp][(w.qn & 0xffff) + __popc(mf & w.lt_mask)] = make_float4(fx0, fy0, fx1, fy1);
    }
    w.qn += __popc(mf);
#endif
}
__device__ __forceinline__ int queue_fill(int qn) { return (qn & 0xffff) + (qn >> 16); }
// make room for `room` more lines
__device__ __forceinline__ void ensure_room(Warp& w, int room, const Canvas& cv) {
    if (queue_fill(w.qn) > kGQueue - room) {
        __syncwarp();
        accumulate_warp(w.qn, cv);
        w.qn = 0;
    }
}

// Stack-free depth-first walk below `root` in the reference's order (left half first): the next node is recomputed
// from the root along its path bits.  Curves on the f64 path and nodes still not flat three levels below a slot root get
// here; their leaves are rasterized by this lane.  `chk`: test every node for NaN like the reference (src/path.rs:765).
__device__ __noinline__ uint32_t walk_deep(const Nd root, const int kind, const bool chk, const double thr, const Canvas cv,
                                           Status* __restrict__ status) {
    Nd cur = root;
    int depth = 0;
    uint32_t path = 0, count = 0;
    while (true) {
        if (chk && nd_has_nan(cur, kind)) { atomicExch(&status->nan_flag, 1u); break; }
        if (nd_flatness(cur, kind) < thr) {
            count++;
            line_slow(cur.x0, cur.y0, nd_endx(cur, kind), nd_endy(cur, kind), cv);
            while (depth > 0 && (path & 1u)) { path >>= 1; depth--; }
            if (depth == 0) break;
            path |= 1u;
            cur = root;
#pragma unroll 1
            for (int lv = depth - 1; lv >= 0; lv--) nd_child(cur, kind, ((path >> lv) & 1u) != 0);
        } else if (depth >= kGMaxBelow) {
            atomicExch(&status->depth_flag, 1u);
            break;
        } else {
            nd_child(cur, kind, false);
            depth++;
            path <<= 1;
        }
    }
    return count;
}

// Expand the queued deep nodes, one lane per child (batches of 16 nodes) while more than `keep` are queued.  All of
// them belong to safe curves.  Returns the new fill level of the line queue; *lines_add: leaves found by this lane.
__device__ __noinline__ int drain_deep(int qn, int dn, const int keep, uint32_t* lines_add, const double thr, const Canvas cv,
                                       Status* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    Warp w;
    w.qn = qn;
    w.dn = 0;
    w.lines = 0;
    w.lt_mask = (1u << lane) - 1u;
    w.warp = threadIdx.x >> 5;
    const double* dq = s_dq[w.warp];
    __syncwarp();
    while (dn > keep) {
        const int take = min(dn, 16);
        dn -= take;
        const bool valid = lane < 2 * take;
        Nd c;
        int kind = 4;
        c.x0 = c.x1 = c.x2 = c.x3 = c.y0 = c.y1 = c.y2 = c.y3 = 0.0;
        if (valid) {
            const int e = dn + (lane >> 1);
            const double2* src = reinterpret_cast<const double2*>(dq + 8 * e);
            const double2 a = src[0], b = src[1], cc = src[2], dd = src[3];
            c.x0 = a.x; c.x1 = a.y; c.x2 = b.x; c.x3 = b.y; c.y0 = cc.x; c.y1 = cc.y; c.y2 = dd.x; c.y3 = dd.y;
            kind = s_dkind[w.warp][e];
            nd_child(c, kind, (lane & 1) != 0);
        }
        const bool flat = valid && nd_flatness(c, kind) < thr;
        ensure_room(w, 32, cv);
        emit_site(flat, c.x0, c.y0, nd_endx(c, kind), nd_endy(c, kind), w);
        if (valid && !flat) w.lines += walk_deep(c, kind, false, thr, cv, status);
    }
    *lines_add = w.lines;
    return w.qn;
}
// nodes left in the deep queue after drain_deep(.., dn, keep, ..)
__device__ __forceinline__ int deep_left(int dn, int keep) {
    while (dn > keep) dn -= min(dn, 16);
    return dn;
}

// Warp-collective: lanes with `pred` hold a node that is not flat two levels below its slot root.  No call in here: when
// the queue cannot take the site's nodes (it is drained at the drain points, so this needs half the lanes to overflow at
// once), the lanes note the site in `ovf` and redo those nodes after the walk.
template <int SITE>
__device__ __forceinline__ void deep_site(bool pred, const Nd& n, int kind, Warp& w, unsigned& ovf) {
    const unsigned m = __ballot_sync(kFull, pred);
    if (m == 0) return;
    if (w.dn + __popc(m) > kGDeep) {
        if (pred) ovf |= 1u << SITE;
        return;
    }
    if (pred) {
        const int e = w.dn + __popc(m & w.lt_mask);
        double2* dst = reinterpret_cast<double2*>(s_dq[w.warp] + 8 * e);
        dst[0] = make_double2(n.x0, n.x1);
        dst[1] = make_double2(n.x2, n.x3);
        dst[2] = make_double2(n.y0, n.y1);
        dst[3] = make_double2(n.y2, n.y3);
        s_dkind[w.warp][e] = (unsigned char)kind;
    }
    w.dn += __popc(m);
}

__device__ __forceinline__ void drain_deep_keep(Warp& w, int keep, const double thr, const Canvas& cv, Status* __restrict__ status) {
    uint32_t add = 0;
    w.qn = drain_deep(w.qn, w.dn, keep, &add, thr, cv, status);
    w.dn = deep_left(w.dn, keep);
    w.lines += add;
}

// One round of a warp: lane = (curve, slot).  `has`: this lane owns a slot; (rx, ry): the curve's transformed control
// points in shared memory (4 doubles each); meta = kind | 8 (f64 path).  Warp-collective.
__device__ __forceinline__ void walk_slot(const bool has, const double* rx, const double* ry, const int meta, const uint32_t slot, Warp& w,
                                          const Canvas& cv, const double thr, Status* __restrict__ status) {
        bool act = false, leaf = false, slow = false;
        int kind = 4;
        Nd nd;
        nd.x0 = nd.x1 = nd.x2 = nd.x3 = nd.y0 = nd.y1 = nd.y2 = nd.y3 = 0.0;
        if (has) {
            kind = meta & 7;
            slow = (meta & 8) != 0;
            const double2 xa = *reinterpret_cast<const double2*>(rx);
            const double2 xb = *reinterpret_cast<const double2*>(rx + 2);
            const double2 ya = *reinterpret_cast<const double2*>(ry);
            const double2 yb = *reinterpret_cast<const double2*>(ry + 2);
            nd.x0 = xa.x; nd.x1 = xa.y; nd.x2 = xb.x; nd.x3 = xb.y;
            nd.y0 = ya.x; nd.y1 = ya.y; nd.y2 = yb.x; nd.y3 = yb.y;
            act = true;
            // descend to this slot's subtree root (bits of `slot`, most significant first).  A NaN control point makes every
            // flatness NaN (never < thr): such a curve descends to the cut and is reported by walk_deep below.
#pragma unroll 1
            for (int level = 0; level < kGDepth; level++) {
                if (nd_flatness(nd, kind) < thr) {
                    // a leaf above the cut: owned by the slot whose remaining bits are all zero
                    if (slot & ((1u << (kGDepth - level)) - 1u)) act = false;
                    else leaf = true;
                    break;
                }
                nd_child(nd, kind, ((slot >> (kGDepth - 1 - level)) & 1u) != 0);
            }
        }
        slow = slow && act;  // this slot's whole subtree on the f64 path, after the walk
        act = act && !slow;
        // queue budget of the first half of the walk: the slot root or its first child, then that child's two children
        ensure_room(w, 96, cv);
        if (act && !leaf) leaf = nd_flatness(nd, kind) < thr;
        // The two levels below the slot root in straight-line code without a call: every child's flatness is tested as
        // soon as it exists.  Queue budget: a lane emits at most three lines in the first half of the walk and two in the second.
        emit_site(act && leaf, nd.x0, nd.y0, nd_endx(nd, kind), nd_endy(nd, kind), w);
        const bool go = act && !leaf;
        unsigned ovf = 0;
        if (__any_sync(kFull, go)) {
            // one child at a time (the sibling is recomputed from its parent: 14 more f64 operations per split than
            // splitting once, but only three nodes are ever live)
            auto leaf2 = [&](const Nd& b, const bool ga, auto site) {
                const bool fb = ga && nd_flatness(b, kind) < thr;
                emit_site(fb, b.x0, b.y0, nd_endx(b, kind), nd_endy(b, kind), w);
                deep_site<decltype(site)::value>(ga && !fb, b, kind, w, ovf);
            };
            auto level1 = [&](const Nd& a, auto side) {
                const bool fa = go && nd_flatness(a, kind) < thr;
                emit_site(fa, a.x0, a.y0, nd_endx(a, kind), nd_endy(a, kind), w);
                const bool ga = go && !fa;
                if (__any_sync(kFull, ga)) {
                    leaf2(nd_half<false>(a, kind), ga, std::integral_constant<int, decltype(side)::value * 2>{});
                    leaf2(nd_half<true>(a, kind), ga, std::integral_constant<int, decltype(side)::value * 2 + 1>{});
                }
            };
            level1(nd_half<false>(nd, kind), std::integral_constant<int, 0>{});
            // drain point: only the slot root is live
            if (w.dn >= 16) drain_deep_keep(w, 15, thr, cv, status);
            ensure_room(w, 64, cv);
            level1(nd_half<true>(nd, kind), std::integral_constant<int, 1>{});
        }
        if (w.dn >= 16) drain_deep_keep(w, 15, thr, cv, status);
        // The round's lines stay in the queue: the leaves of the deep nodes and the line items join them, and the kernel drains
        // everything in one pass at the end (fewer, fuller groups of 32).  The next round makes room first (ensure_room above).
        // rare: nodes the deep queue could not take, and slots of curves on the f64 path
        if (ovf) {
            for (int site = 0; site < 4; site++)
                if ((ovf >> site) & 1u) {
                    Nd b = nd;
                    nd_child(b, kind, (site >> 1) != 0);
                    nd_child(b, kind, (site & 1) != 0);
                    w.lines += walk_deep(b, kind, false, thr, cv, status);
                }
        }
        if (slow) {
            if (leaf) {
                w.lines++;
                if (nd_has_nan(nd, kind)) atomicExch(&status->nan_flag, 1u);
                else line_slow(nd.x0, nd.y0, nd_endx(nd, kind), nd_endy(nd, kind), cv);
            } else {
                w.lines += walk_deep(nd, kind, true, thr, cv, status);
            }
        }
    }

// One line / closing item per lane (items beyond the curves in the packed order).  Warp-collective.
__device__ __forceinline__ void emit_line_item(const uint32_t i, const uint32_t n_items, const JobDev& job, Warp& w, const Canvas& cv,
                                               Status* __restrict__ status) {
        bool pred = false;
        double x0 = 0.0, y0 = 0.0, x1 = 0.0, y1 = 0.0;
        if (i < n_items) {
            const uint2 it = job.items_packed[i];
            uint32_t ia, ib;
            if (it.y & kItemClosing) {
                // Line::new(subpath.end(), subpath.start()).transform(tr), src/path.rs:781-785: emitted when the subpath
                // is closed or `close` is set, even if zero length
                pred = (it.y & kItemExplicitClosed) || job.close;
                ia = it.x;
                ib = it.y & kItemIndexMask;
            } else {
                pred = true;
                ia = it.x;
                ib = it.x + 1;
            }
            if (pred) {
                const P2 a = tr_apply(job.tr, job.pts[ia]);
                const P2 b = tr_apply(job.tr, job.pts[ib]);
                x0 = a.x; y0 = a.y; x1 = b.x; y1 = b.y;
                // Line: flatness 0 < thr (src/curve.rs:233-235); NaN end points panic in the reference (src/path.rs:765-767)
                if (isnan(x0) || isnan(y0) || isnan(x1) || isnan(y1)) {
                    atomicExch(&status->nan_flag, 1u);
                    pred = false;
                } else if (!(point_safe(x0, y0, cv) && point_safe(x1, y1, cv))) {
                    w.lines++;
                    line_slow(x0, y0, x1, y1, cv);
                    pred = false;
                }
            }
        }
        ensure_room(w, 32, cv);
        emit_site(pred, x0, y0, x1, y1, w);
}

// K3 phase 2 + K4 for the rows of NWARPS warps (warp = 0 .. NWARPS - 1, tid = thread index among them): row scan, fill
// rule, composite, store.  `cells`: the canvas' cell buffer; `paint_buf`: shared memory for a gradient paint's table.
template <bool PLAIN, int NWARPS>
__device__ __forceinline__ void finish_rows(const JobDev& job, const PaintDev* __restrict__ paints, int* const cells, void* paint_buf, const int warp,
                                            const int lane, const int tid, Status* __restrict__ status) {
    const Fix fix = make_fix(job.fix_shift);
    // ---- K3 phase 2 + K4: two rows per warp iteration (16 lanes x 4 columns each) ----------------------------
    const int mode = job.mode;
    const bool render = mode == kModeRender;  // fill onto a canvas created here: every pixel is written, none is read
    const int wout = job.width_out, hout = job.height;
    const int half = lane >> 4, hl = lane & 15;
    const bool evenodd = job.rule == 1;
    const bool has_paint = mode >= kModeFill && job.paint_index >= 0;
    const PaintDev& s_paint = *reinterpret_cast<const PaintDev*>(paint_buf);
    bool solid = true;
    float4 solid_c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_paint) {
        const PaintDev* gp = &paints[job.paint_index];
        solid = PLAIN || gp->kind == 0;
        if (solid) {
            solid_c = *reinterpret_cast<const float4*>(gp->solid);
        } else {  // gradient: its table replaces the line queues
            const int* src = reinterpret_cast<const int*>(gp);
            int* dst = reinterpret_cast<int*>(paint_buf);
            for (int i = tid; i < (int)(sizeof(PaintDev) / 4); i += NWARPS * 32) dst[i] = src[i];
            if (NWARPS == 1) __syncwarp(); else __syncthreads();
        }
    }
    // RENDER with a solid colour whose components are finite and not negative (the glyph batch): the composite
    // `0.blend_over(colour * alpha)` is colour * alpha exactly — colour * alpha + 0 * (1 - colour.a * alpha) adds +0 to a
    // value that is never -0 — so a pixel is four multiplications and one store
    const bool plain = render && solid && solid_c.x >= 0.f && solid_c.y >= 0.f && solid_c.z >= 0.f && solid_c.w >= 0.f &&
                       !signbit(solid_c.x) && !signbit(solid_c.y) && !signbit(solid_c.z) && !signbit(solid_c.w) && solid_c.x < 3e38f &&
                       solid_c.y < 3e38f && solid_c.z < 3e38f && solid_c.w < 3e38f;
    float4* const out_base = reinterpret_cast<float4*>(job.canvas) + job.origin;
    const unsigned long long row_stride = job.row_stride;
    for (int r2 = warp * 2; r2 < hout; r2 += NWARPS * 2) {
        const int r = r2 + half;
        const bool rvalid = r < hout;
        int* rowc = cells + (rvalid ? r : 0) * kSmPitch;
        int4 qv = make_int4(0, 0, 0, 0);
        if (rvalid) qv = *reinterpret_cast<const int4*>(rowc + hl * 4);
        const int p0 = qv.x, p1 = p0 + qv.y, p2 = p1 + qv.z, p3 = p2 + qv.w;
        int incl = p3;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int nb = __shfl_up_sync(kFull, incl, o, 16);
            if (hl >= o) incl += nb;
        }
        const int base = incl - p3;
        float4 c;
        if (evenodd) {
            c = make_float4(coverage_from_fixed<true>(base + p0, fix), coverage_from_fixed<true>(base + p1, fix), coverage_from_fixed<true>(base + p2, fix),
                            coverage_from_fixed<true>(base + p3, fix));
        } else {
            if (winding_risk(base + p0, base + p1, base + p2, base + p3, fix)) status->winding_flag = 1u;  // see rgpu_internal.cuh: kFixShift
            c = make_float4(coverage_from_fixed<false>(base + p0, fix), coverage_from_fixed<false>(base + p1, fix),
                            coverage_from_fixed<false>(base + p2, fix), coverage_from_fixed<false>(base + p3, fix));
        }
        const int col = hl * 4;
        if (mode != kModeMask) {  // mask_iter drops abs(alpha) < 1e-6 (src/rasterize.rs:348); fill_impl sees that iterator
            if (c.x < 1e-6f) c.x = 0.f;
            if (c.y < 1e-6f) c.y = 0.f;
            if (c.z < 1e-6f) c.z = 0.f;
            if (c.w < 1e-6f) c.w = 0.f;
        }
        if (mode < kModeFill) {
            if (rvalid) {
                float* out = reinterpret_cast<float*>(job.canvas) + job.origin + (unsigned long long)r * job.row_stride;
                if (col + 3 < wout && ((reinterpret_cast<uintptr_t>(out + col) & 15) == 0)) {
                    __stcs(reinterpret_cast<float4*>(out + col), c);
                } else {
                    if (col < wout) out[col] = c.x;
                    if (col + 1 < wout) out[col + 1] = c.y;
                    if (col + 2 < wout) out[col + 2] = c.z;
                    if (col + 3 < wout) out[col + 3] = c.w;
                }
            }
            continue;
        }
        // stage the two rows' coverage so that consecutive lanes composite consecutive pixels (16 B each).  (Every lane storing
        // the four pixels it holds instead — 64 B per lane, no staging — was measured: 3.92 vs 3.51 ms per 100 000 glyphs, the
        // strided 16-byte stores cost more than the shared-memory round trip.)
        if (rvalid) *reinterpret_cast<float4*>(rowc + col) = c;
        __syncwarp();
        if (plain && wout == 64 && r2 + 1 < hout) {
            const float* al = reinterpret_cast<const float*>(cells + r2 * kSmPitch);
            float4* o0 = out_base + (unsigned long long)r2 * row_stride;
            float4* o1 = o0 + row_stride;
            const float a0 = al[lane], a1 = al[lane + 32], a2 = al[kSmPitch + lane], a3 = al[kSmPitch + lane + 32];
            __stcs(o0 + lane, make_float4(fmul(solid_c.x, a0), fmul(solid_c.y, a0), fmul(solid_c.z, a0), fmul(solid_c.w, a0)));
            __stcs(o0 + lane + 32, make_float4(fmul(solid_c.x, a1), fmul(solid_c.y, a1), fmul(solid_c.z, a1), fmul(solid_c.w, a1)));
            __stcs(o1 + lane, make_float4(fmul(solid_c.x, a2), fmul(solid_c.y, a2), fmul(solid_c.z, a2), fmul(solid_c.w, a2)));
            __stcs(o1 + lane + 32, make_float4(fmul(solid_c.x, a3), fmul(solid_c.y, a3), fmul(solid_c.z, a3), fmul(solid_c.w, a3)));
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int pidx = i * 32 + lane;  // 0..127 over the two rows
                const int rr = r2 + (pidx >> 6), px = pidx & 63;
                if (rr < hout && px < wout) {
                    const float alpha = reinterpret_cast<const float*>(cells + rr * kSmPitch)[px];
                    float4* out = out_base + (unsigned long long)rr * row_stride;
                    if (alpha != 0.0f) {
                        float4 color = solid_c;
                        if (!PLAIN && !solid) color = paint_at(s_paint, px, rr);
                        color = make_float4(fmul(color.x, alpha), fmul(color.y, alpha), fmul(color.z, alpha), fmul(color.w, alpha));
                        float4 dstc = render ? make_float4(0.f, 0.f, 0.f, 0.f) : out[px];  // `Layer::new`: transparent
                        const float k = fsub(1.0f, color.w);
                        dstc = make_float4(fadd(color.x, fmul(dstc.x, k)), fadd(color.y, fmul(dstc.y, k)), fadd(color.z, fmul(dstc.z, k)),
                                           fadd(color.w, fmul(dstc.w, k)));
                        if (render) __stcs(out + px, dstc); else out[px] = dstc;
                    } else if (render) {
                        __stcs(out + px, make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                }
            }
        }
        __syncwarp();
    }
}

// PLAIN: no gradient paint in the batch (masks, coverage, solid fills) — the paint evaluation is compiled out
template <int MINB, bool PLAIN>
__global__ void __launch_bounds__(kGThreads, MINB)
small_canvas_kernel(const JobDev* __restrict__ jobs, uint32_t job_first, const PaintDev* __restrict__ paints, double thr,
                    Status* __restrict__ status) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    {
        const int* src = reinterpret_cast<const int*>(&jobs[job_first + blockIdx.x]);
        int* dst = reinterpret_cast<int*>(&s_job);
        if (tid < (int)(sizeof(JobDev) / 4)) dst[tid] = src[tid];
        const int4 z = make_int4(0, 0, 0, 0);
        int4* c4 = reinterpret_cast<int4*>(s_cells);
        for (int i = tid; i < kCells / 4; i += kGThreads) c4[i] = z;
        if (tid == 0) s_nlines = 0;
    }
    __syncthreads();
    const JobDev& job = s_job;
    Canvas cv;
    cv.H = job.height;
    cv.wc = job.clamp_w;
    cv.wci = (int)cv.wc;
    cv.wcf = (float)cv.wc;
    cv.tile_end = min(kSmPitch, cv.wci + 1);  // reference columns incl. the overflow column
    cv.fix_scale = make_fix(job.fix_shift).scale;
    Warp w;
    w.qn = 0;
    w.dn = 0;
    w.lines = 0;
    w.lt_mask = (1u << lane) - 1u;
    w.warp = warp;
    const uint32_t n_items = job.n_items, n_curves = job.n_curves;
    // ---- curves (first in the packed order): chunks of 20, 8 slot threads each --------------------------------
    for (uint32_t c0 = 0; c0 < n_curves; c0 += kGCurves) {
        const uint32_t cn = min((uint32_t)kGCurves, n_curves - c0);
        if (c0) __syncthreads();  // the previous chunk's slots are done with s_x / s_y
        if (tid < kGCurves * 4) {
            // stage: thread = (curve, point); Transform::apply once per control point
            const uint32_t i = (uint32_t)tid >> 2, j = (uint32_t)tid & 3u;
            bool slow = false;
            int kind = 0;
            if (i < cn) {
                const uint2 it = job.items_packed[c0 + i];
                kind = (int)it.y;
                if (j < (uint32_t)kind) {
                    const P2 p = tr_apply(job.tr, job.pts[it.x + j]);
                    s_x[i][j] = p.x;
                    s_y[i][j] = p.y;
                    slow = !point_safe(p.x, p.y, cv);  // also true for NaN / infinite coordinates
                } else {
                    s_x[i][j] = 0.0;
                    s_y[i][j] = 0.0;
                }
            }
            // kGCurves * 4 = 80 threads: warps 0, 1 and the lower half of warp 2 — the four lanes of a curve are in one warp.
            // (no short-circuit around the shuffles: every lane of the group must execute them)
            const unsigned grp = __activemask();
            const bool s1 = __shfl_xor_sync(grp, slow, 1);
            slow = slow | s1;
            const bool s2 = __shfl_xor_sync(grp, slow, 2);
            slow = slow | s2;
            if (i < cn && j == 0) s_meta[i] = (unsigned char)(kind | (slow ? 8 : 0));
        }
        __syncthreads();

        // one round: thread = (curve, slot).  The chunk's slots are dealt out evenly over the warps (a glyph's 18 curves are
        // 144 slots: 29 per warp instead of 32, 32, 32, 32, 16), so that the warps reach the barrier before the row scan together
        const uint32_t n_slots = cn << kGDepth;
        const uint32_t per_warp = (n_slots + kGWarps - 1) / kGWarps;
#ifdef RGPU_SLOTS_CONTIGUOUS
        const uint32_t sidx = (uint32_t)warp * per_warp + (uint32_t)lane;
#else
        // interleaved: warp w takes slots w, w + 5, w + 10, ... — the 8 slots of a curve (and with them its depth and its
        // lines) are spread over all warps instead of making one warp's share
        const uint32_t sidx = (uint32_t)lane * kGWarps + (uint32_t)warp;
#endif
        const bool has_slot = (uint32_t)lane < per_warp && sidx < n_slots;
        const uint32_t ci = has_slot ? (sidx >> kGDepth) : cn, slot = sidx & (kGSlots - 1);
        walk_slot(ci < cn, &s_x[min(ci, (uint32_t)kGCurves - 1)][0], &s_y[min(ci, (uint32_t)kGCurves - 1)][0], ci < cn ? s_meta[ci] : 0, slot, w, cv, thr, status);
    }
    // ---- lines and closing lines: one thread each, straight from the path -------------------------------------
    for (uint32_t i0 = n_curves; i0 < n_items; i0 += kGThreads) {
        const uint32_t i = i0 + tid;
        emit_line_item(i, n_items, job, w, cv, status);
    }
    if (w.dn) drain_deep_keep(w, 0, thr, cv, status);
    if (w.qn) {
        __syncwarp();
        accumulate_warp(w.qn, cv);
        w.qn = 0;
    }
    {
        uint32_t v = w.lines;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0 && v) atomicAdd(&s_nlines, v);
    }
    __syncthreads();
    if (tid == 0 && s_nlines) atomicAdd(&status->n_lines, s_nlines);  // statistics only

    finish_rows<PLAIN, kGWarps>(job, paints, s_cells, &s_lineq[0][0], warp, lane, tid, status);
}


}  // namespace

bool small_canvas_eligible(uint32_t width, uint32_t height, int mode) {
    // COVERAGE / FILL use one extra (overflow) column inside the tile
    const uint32_t cols = (mode == kModeMask) ? width : width + 1;
    return cols <= (uint32_t)kSmMaxW + 1 && width <= (uint32_t)kSmMaxW && height <= (uint32_t)kSmMaxH;
}

void launch_small_canvas(const JobDev* jobs, uint32_t job_first, uint32_t n_jobs, const PaintDev* paints, double thr, Status* status,
                         bool gradients, cudaStream_t s) {
    if (n_jobs == 0) return;
    // CTAs per SM (register budget): 4 -> 96 registers, 5 -> 72 with spills in the walk (measured: 0.94 vs 1.02 ms per 20 000 glyphs)
    static const int minb = getenv("RGPU_SMALL_MINB") ? atoi(getenv("RGPU_SMALL_MINB")) : 4;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(small_canvas_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        cudaFuncSetAttribute(small_canvas_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        cudaFuncSetAttribute(small_canvas_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        configured[dev] = true;
    }
    if (gradients) {
        small_canvas_kernel<4, false><<<n_jobs, kGThreads, kSmemBytes, s>>>(jobs, job_first, paints, thr, status);
    } else if (minb <= 4) {
        small_canvas_kernel<4, true><<<n_jobs, kGThreads, kSmemBytes, s>>>(jobs, job_first, paints, thr, status);
    } else {
        small_canvas_kernel<5, true><<<n_jobs, kGThreads, kSmemBytes, s>>>(jobs, job_first, paints, thr, status);
    }
}

}  // namespace rgpu
