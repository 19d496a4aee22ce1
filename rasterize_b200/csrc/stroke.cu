// `Path::stroke` on the device (reference src/path.rs:374-415, 692-732): the outline of a stroked path as an ordinary
// device-resident path, ready for K1 (flatten).  Compiled with --fmad=false (see stroke_device.cuh).
//
// The reference walks the subpaths serially, appending to one segment list: every source segment is offset by
// width / 2 (curves recursively, up to 8 pieces and the round joins between them), a join (or, at the turning point of
// an open subpath, a cap) is inserted in front of its pieces when something was emitted before it, and the contour is
// finished by `stroke_close` (closed subpaths: once forward, once backward) or by the final cap (open subpaths).
// Here one thread owns one **unit** of that walk:
//   * an ordinary unit = (source segment, direction): its pieces depend on nothing else; the join in front of them needs
//     the LAST piece of the nearest earlier unit of the contour that emitted anything and its own FIRST piece;
//   * a closer unit per contour = what `stroke_close` / the final cap append, from the contour's first and last piece.
// Three passes, no atomics, output in the reference's order:
//   pieces  (count the pieces of every ordinary unit, keep its first and last piece)
//   count   (add the join / closer segments to the counts) -> exclusive scans of segments, points, curves, contours
//   emit    (recompute and write points + item lists at the scanned offsets).
// The unit table is built on the host from the path's structure (kinds, subpath offsets, closed flags) alone
// (stroke_units.hpp); the per-unit code is in stroke_device.cuh.
#include "rgpu_internal.cuh"
#include "stroke_device.cuh"

namespace rgpu {

using namespace sk;

namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) stroke_pieces_kernel(const StrokeUnit* __restrict__ units, uint32_t n_units,
                                                                 const double* __restrict__ pts, Style style, uint32_t* __restrict__ cnt,
                                                                 PieceRec* __restrict__ first, PieceRec* __restrict__ last) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_units) return;
    const uint32_t stride = stroke_count_stride(n_units);
    unit_pieces(i, units, pts, style, cnt, cnt + stride, cnt + 2 * stride, first, last);
}

// cnt is read (other units' segment counts, as "has pieces" flags) and written (this unit's) by the count pass
__global__ void __launch_bounds__(kThreads) stroke_count_kernel(const StrokeUnit* __restrict__ units, uint32_t n_units,
                                                                const double* __restrict__ pts, Style style, uint32_t* cnt,
                                                                const PieceRec* __restrict__ first, const PieceRec* __restrict__ last) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_units) return;
    const uint32_t stride = stroke_count_stride(n_units);
    unit_count(i, units, n_units, pts, style, cnt, cnt + stride, cnt + 2 * stride, cnt + 3 * stride, first, last);
}

__global__ void __launch_bounds__(kThreads) stroke_emit_kernel(const StrokeUnit* __restrict__ units, uint32_t n_units,
                                                               const double* __restrict__ pts, Style style, const uint32_t* __restrict__ cnt,
                                                               const PieceRec* __restrict__ first, const PieceRec* __restrict__ last,
                                                               const uint32_t* __restrict__ off, double* out_pts, uint2* out_items,
                                                               uint2* out_packed) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_units) return;
    const uint32_t stride = stroke_count_stride(n_units);
    const uint32_t* off_seg = off;
    const uint32_t* off_pts = off + stride;
    const uint32_t* off_curves = off + 2 * stride;
    const uint32_t* off_close = off + 3 * stride;
    EmitSink<uint2> sink;
    sink.pts = out_pts;
    sink.items = out_items;
    sink.packed = out_packed;
    sink.pt = off_pts[i];
    sink.item = off_seg[i] + off_close[i];
    sink.curve = off_curves[i];
    sink.total_curves = off_curves[n_units];  // exclusive scan over n_units + 1 entries: the last one is the total
    unit_emit(i, units, n_units, pts, style, cnt, first, last, off_pts[units[i].c], kItemClosing | kItemExplicitClosed, sink);
}

Style to_style(const StrokeStyleDev& st) {
    Style style;
    style.width = st.width;
    style.miter_limit = st.miter_limit;
    style.join = st.join;
    style.cap = st.cap;
    return style;
}

}  // namespace

void launch_stroke_pieces(const StrokeUnit* units, uint32_t n_units, const double2* pts, const StrokeStyleDev& st, uint32_t* cnt, void* first,
                          void* last, cudaStream_t s) {
    const dim3 grid((n_units + kThreads - 1) / kThreads);
    stroke_pieces_kernel<<<grid, kThreads, 0, s>>>(units, n_units, reinterpret_cast<const double*>(pts), to_style(st), cnt,
                                                   static_cast<PieceRec*>(first), static_cast<PieceRec*>(last));
}

void launch_stroke_units(bool emit, const StrokeUnit* units, uint32_t n_units, const double2* pts, const StrokeStyleDev& st, uint32_t* cnt,
                         const void* first, const void* last, const uint32_t* off, double2* out_pts, uint2* out_items, uint2* out_packed,
                         cudaStream_t s) {
    const dim3 grid((n_units + kThreads - 1) / kThreads);
    const PieceRec* f = static_cast<const PieceRec*>(first);
    const PieceRec* l = static_cast<const PieceRec*>(last);
    const double* p = reinterpret_cast<const double*>(pts);
    if (emit)
        stroke_emit_kernel<<<grid, kThreads, 0, s>>>(units, n_units, p, to_style(st), cnt, f, l, off, reinterpret_cast<double*>(out_pts), out_items,
                                                     out_packed);
    else
        stroke_count_kernel<<<grid, kThreads, 0, s>>>(units, n_units, p, to_style(st), cnt, f, l);
}

size_t stroke_piece_bytes() { return sizeof(PieceRec); }

}  // namespace rgpu
