// Device-side curve flattening primitives shared by flatten.cu (global line buffers) and small.cu (lines kept in
// shared memory).  See flatten.cu for the reference citations and the bit-exactness argument.
#pragma once
#include "rgpu_internal.cuh"

namespace rgpu {
namespace fl {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

struct P2 { double x, y; };

// Transform::apply: x*m00 + y*m01 + m02 (two rounded products, two rounded sums)
__device__ __forceinline__ P2 tr_apply(const double* m, double2 p) {
    P2 r;
    r.x = dadd(dadd(dmul(p.x, m[0]), dmul(p.y, m[1])), m[2]);
    r.y = dadd(dadd(dmul(p.x, m[3]), dmul(p.y, m[4])), m[5]);
    return r;
}

struct Seg {
    P2 p[4];
};

__device__ __forceinline__ bool seg_has_nan(const Seg& s, int kind) {
    bool n = false;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < kind) n = n || isnan(s.p[i].x) || isnan(s.p[i].y);
    return n;
}

// Curve::end(): selected without a runtime array index, which would push the whole Seg into local memory
__device__ __forceinline__ P2 seg_end(const Seg& s, int kind) { return kind == 4 ? s.p[3] : (kind == 3 ? s.p[2] : s.p[1]); }

// Rust f64::max: NaN operands are ruled out by has_nans before this is evaluated
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

__device__ __forceinline__ double seg_flatness(const Seg& s, int kind) {
    if (kind == 2) return 0.0;
    if (kind == 3) {
        // 2.0 * p1 - p0 - p2
        double dx = dsub(dsub(dmul(2.0, s.p[1].x), s.p[0].x), s.p[2].x);
        double dy = dsub(dsub(dmul(2.0, s.p[1].y), s.p[0].y), s.p[2].y);
        return dadd(dmul(dx, dx), dmul(dy, dy));
    }
    // u = 3.0 * p1 - 2.0 * p0 - p3 ; v = 3.0 * p2 - p0 - 2.0 * p3
    double ux = dsub(dsub(dmul(3.0, s.p[1].x), dmul(2.0, s.p[0].x)), s.p[3].x);
    double uy = dsub(dsub(dmul(3.0, s.p[1].y), dmul(2.0, s.p[0].y)), s.p[3].y);
    double vx = dsub(dsub(dmul(3.0, s.p[2].x), s.p[0].x), dmul(2.0, s.p[3].x));
    double vy = dsub(dsub(dmul(3.0, s.p[2].y), s.p[0].y), dmul(2.0, s.p[3].y));
    return dadd(dmax(dmul(ux, ux), dmul(vx, vx)), dmax(dmul(uy, uy), dmul(vy, vy)));
}

__device__ __forceinline__ double lerp3(double a, double ca, double b, double cb, double c, double cc) {
    return dadd(dadd(dmul(ca, a), dmul(cb, b)), dmul(cc, c));
}

// split(): s0 = left half, s1 = right half
__device__ __forceinline__ void seg_split(const Seg& s, int kind, Seg& s0, Seg& s1) {
    if (kind == 4) {
        P2 mid;
        mid.x = dadd(dadd(dadd(dmul(0.125, s.p[0].x), dmul(0.375, s.p[1].x)), dmul(0.375, s.p[2].x)), dmul(0.125, s.p[3].x));
        mid.y = dadd(dadd(dadd(dmul(0.125, s.p[0].y), dmul(0.375, s.p[1].y)), dmul(0.375, s.p[2].y)), dmul(0.125, s.p[3].y));
        s0.p[0] = s.p[0];
        s0.p[1].x = dadd(dmul(0.5, s.p[0].x), dmul(0.5, s.p[1].x));
        s0.p[1].y = dadd(dmul(0.5, s.p[0].y), dmul(0.5, s.p[1].y));
        s0.p[2].x = lerp3(s.p[0].x, 0.25, s.p[1].x, 0.5, s.p[2].x, 0.25);
        s0.p[2].y = lerp3(s.p[0].y, 0.25, s.p[1].y, 0.5, s.p[2].y, 0.25);
        s0.p[3] = mid;
        s1.p[0] = mid;
        s1.p[1].x = lerp3(s.p[1].x, 0.25, s.p[2].x, 0.5, s.p[3].x, 0.25);
        s1.p[1].y = lerp3(s.p[1].y, 0.25, s.p[2].y, 0.5, s.p[3].y, 0.25);
        s1.p[2].x = dadd(dmul(0.5, s.p[2].x), dmul(0.5, s.p[3].x));
        s1.p[2].y = dadd(dmul(0.5, s.p[2].y), dmul(0.5, s.p[3].y));
        s1.p[3] = s.p[3];
    } else if (kind == 3) {
        // mid = 0.25 * (p0 + 2.0 * p1 + p2)
        P2 mid;
        mid.x = dmul(0.25, dadd(dadd(s.p[0].x, dmul(2.0, s.p[1].x)), s.p[2].x));
        mid.y = dmul(0.25, dadd(dadd(s.p[0].y, dmul(2.0, s.p[1].y)), s.p[2].y));
        s0.p[0] = s.p[0];
        s0.p[1].x = dmul(0.5, dadd(s.p[0].x, s.p[1].x));
        s0.p[1].y = dmul(0.5, dadd(s.p[0].y, s.p[1].y));
        s0.p[2] = mid;
        s1.p[0] = mid;
        s1.p[1].x = dmul(0.5, dadd(s.p[1].x, s.p[2].x));
        s1.p[1].y = dmul(0.5, dadd(s.p[1].y, s.p[2].y));
        s1.p[2] = s.p[2];
    } else {
        // Line: default split_at(0.5): mid = (1.0 - t) * p0 + t * p1 (never reached: flatness 0 < thr)
        P2 mid;
        mid.x = dadd(dmul(0.5, s.p[0].x), dmul(0.5, s.p[1].x));
        mid.y = dadd(dmul(0.5, s.p[0].y), dmul(0.5, s.p[1].y));
        s0.p[0] = s.p[0]; s0.p[1] = mid;
        s1.p[0] = mid;    s1.p[1] = s.p[1];
    }
}

// State of one (item, slot) after loading, transforming and descending to the slot's subtree root.
struct SlotCtx {
    Seg seg;
    int kind;
    uint32_t job;
    bool leaf_above;  // the subtree root is itself a leaf (the curve was flat above the depth-3 cut)
};

// Loads item `item` of `job`, transforms it and descends to the root of subtree `slot`.  Returns false when the slot
// produces no line.  DEPTH = level at which the subdivision tree is cut into slots.
template <int DEPTH>
__device__ __forceinline__ bool slot_from_item(const JobDev& job, uint32_t j, const uint2 item, uint32_t slot, double thr, SlotCtx& c,
                                               Status* __restrict__ status);

// Thread t of the (item, slot) grid: items in the reference's order, 2^DEPTH slots each.
template <int DEPTH = kSlotDepth>
__device__ __forceinline__ bool slot_setup(const JobDev* __restrict__ jobs, uint32_t n_jobs, uint32_t t, double thr, SlotCtx& c,
                                           Status* __restrict__ status) {
    const uint32_t g = t >> DEPTH;
    const uint32_t slot = t & ((1u << DEPTH) - 1u);
    const uint32_t j = find_job(n_jobs, g, [&](uint32_t k) { return jobs[k].item_begin; });
    const JobDev& job = jobs[j];
    return slot_from_item<DEPTH>(job, j, job.items[g - job.item_begin], slot, thr, c, status);
}

// Thread t of the PACKED grid (raster path): per job, 2^DEPTH slot threads for each of its curves, then ONE thread for
// each line / closing item (`items_packed` holds the curves first).  Lines — most items of a typical path — no longer
// occupy 2^DEPTH threads of which all but one exit at once, so the grid is a fraction of the (item, slot) grid.
template <int DEPTH>
__device__ __forceinline__ bool slot_setup_packed(const JobDev* __restrict__ jobs, uint32_t n_jobs, uint32_t t, double thr, SlotCtx& c,
                                                  Status* __restrict__ status) {
    const uint32_t j = find_job(n_jobs, t, [&](uint32_t k) { return jobs[k].thread_begin; });
    const JobDev& job = jobs[j];
    const uint32_t lt = t - job.thread_begin;
    const uint32_t curve_threads = job.n_curves << DEPTH;
    if (lt < curve_threads) return slot_from_item<DEPTH>(job, j, job.items_packed[lt >> DEPTH], lt & ((1u << DEPTH) - 1u), thr, c, status);
    const uint32_t g = job.n_curves + (lt - curve_threads);
    if (g >= job.n_items) return false;  // padding threads of the job's last warp
    return slot_from_item<DEPTH>(job, j, job.items_packed[g], 0u, thr, c, status);
}

template <int DEPTH>
__device__ __forceinline__ bool slot_from_item(const JobDev& job, uint32_t j, const uint2 item, uint32_t slot, double thr, SlotCtx& c,
                                               Status* __restrict__ status) {
    const double* m = job.tr;
    c.job = j;
    c.leaf_above = false;
    if (item.y & kItemClosing) {
        // closing line of a subpath: Line::new(subpath.end(), subpath.start()).transform(tr), src/path.rs:781-785.
        // Emitted when the subpath is closed or `close` is set, even if zero length.
        const bool emit_it = (item.y & kItemExplicitClosed) || job.close;
        if (slot != 0 || !emit_it) return false;
        c.kind = 2;
        c.seg.p[0] = tr_apply(m, job.pts[item.x]);
        c.seg.p[1] = tr_apply(m, job.pts[item.y & kItemIndexMask]);
    } else {
        c.kind = (int)item.y;
        if (c.kind == 2 && slot != 0) return false;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < c.kind) c.seg.p[i] = tr_apply(m, job.pts[item.x + i]);
    }
    // descend to this slot's subtree root (bits of `slot`, most significant first)
#pragma unroll 1
    for (int level = 0; level < DEPTH; level++) {
        if (seg_has_nan(c.seg, c.kind)) {
            atomicExch(&status->nan_flag, 1u);
            return false;
        }
        if (seg_flatness(c.seg, c.kind) < thr) {
            // a leaf above the cut: owned by the slot whose remaining bits are all zero
            const uint32_t rest = slot & ((1u << (DEPTH - level)) - 1u);
            if (rest != 0) return false;
            c.leaf_above = true;
            return true;
        }
        Seg s0, s1;
        seg_split(c.seg, c.kind, s0, s1);
        c.seg = ((slot >> (DEPTH - 1 - level)) & 1u) ? s1 : s0;
    }
    return true;
}

// Depth-first walk of the slot's subtree in the reference's order (left half first); `stack` holds pending right
// halves.  Calls emit(x0,y0,x1,y1) for every leaf and returns the number of leaves.
// True when every control point is finite and far from overflow: all descendants are convex combinations of the
// root's points, so no NaN can appear below it and the per-node has_nans test of the reference (src/path.rs:765)
// can never fire — slot_walk<false> may then skip it.
__device__ __forceinline__ bool seg_all_finite(const Seg& s, int kind) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < kind) ok = ok && fabs(s.p[i].x) < 1e300 && fabs(s.p[i].y) < 1e300;
    return ok;
}

template <bool CHECK_NAN = true, class Emit>
__device__ __forceinline__ uint32_t slot_walk(const SlotCtx& c, double thr, Status* __restrict__ status, Emit emit) {
    const int kind = c.kind;
    Seg seg = c.seg;
    if (c.leaf_above) {
        const P2 e = seg_end(seg, kind);
        emit(seg.p[0].x, seg.p[0].y, e.x, e.y);
        return 1;
    }
    Seg stack[kMaxStack];
    int top = 0;
    uint32_t count = 0;
    while (true) {
        if (CHECK_NAN && seg_has_nan(seg, kind)) { atomicExch(&status->nan_flag, 1u); break; }
        if (seg_flatness(seg, kind) < thr) {
            const P2 e = seg_end(seg, kind);
            emit(seg.p[0].x, seg.p[0].y, e.x, e.y);
            count++;
            if (top == 0) break;
            seg = stack[--top];
        } else {
            if (top >= kMaxStack) { atomicExch(&status->depth_flag, 1u); break; }
            Seg s0, s1;
            seg_split(seg, kind, s0, s1);
            stack[top++] = s1;
            seg = s0;
        }
    }
    return count;
}

}  // namespace fl
}  // namespace rgpu
