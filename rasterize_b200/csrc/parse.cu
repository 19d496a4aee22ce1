// Batch SVG path parse + `Path::bbox` + `fit_size` on the device (reference src/svg.rs:62-421,
// src/path.rs:428-451, 832-972, src/geometry.rs:470-516).  Compiled with --fmad=false (see parse_device.cuh).
//
// One thread parses one **chunk** of text, start to end: relative commands and `PathBuilder::line_to`'s near-zero test make
// every segment depend on the running position, and an f64 running sum cannot be re-associated without changing bits, so
// text is sequential by construction — up to the next absolute `M`: a moveto finishes the pending subpath and sets every
// piece of parser and builder state from its own operands, so the text from one `M` to the next parses the same alone as
// in place.  A chunk is therefore a whole short string (glyph batches: 1e5 strings, one thread each) or a run of `M ...`
// groups of a long one (the reference's `material-big` parse bench: 3 129 moveto groups in 438 KB), cut on the host with
// memchr.  Two passes with the same parser and a different output policy:
//   count  (segments, closing items, points, curves, bbox, status per chunk; fit_size for single-chunk paths)
//   -> the host adds the chunks of each path up (prefix sums, union of boxes, first error) and plans the output
//   emit   (control points and both item lists at the planned offsets: an ordinary device path batch).
#define SD_FN __host__ __device__
#include <cstring>
#include "rgpu_internal.cuh"
#include "parse_device.cuh"
#include "parse_plan.hpp"

namespace rgpu {

using namespace sv;

namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) parse_count_kernel(const uint8_t* __restrict__ text, const uint32_t* __restrict__ chunk_off,
                                                               uint32_t n_chunks, ParseFit fit, ParseInfoDev* __restrict__ info) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_chunks) return;
    const uint32_t a = chunk_off[i], b = chunk_off[i + 1];
    CountOut out;
    PathBuild<CountOut> builder(out);
    uint32_t err_pos = 0;
    const int status = parse_svg_path(text + a, b - a, builder, err_pos);
    ParseInfoDev r;
    r.status = status;
    r.error_offset = status ? err_pos : 0u;
    const bool ok = status == kParseOk;
    r.n_segments = ok ? out.seg : 0u;
    r.n_subpaths = ok ? out.sub : 0u;
    r.n_points = ok ? out.pts : 0u;
    r.n_curves = ok ? out.curves : 0u;
    r.pad_ = 0;
    r.has_bbox = ok && builder.has_box;
    r.bbox[0] = r.has_bbox ? builder.box.lo.x : 0.0;
    r.bbox[1] = r.has_bbox ? builder.box.lo.y : 0.0;
    r.bbox[2] = r.has_bbox ? builder.box.hi.x : 0.0;
    r.bbox[3] = r.has_bbox ? builder.box.hi.y : 0.0;
    for (int k = 0; k < 6; k++) r.fit_tr[k] = (k == 0 || k == 4) ? 1.0 : 0.0;
    r.fit_width = r.fit_height = 0;
    if (r.has_bbox && fit.align >= 0) fit_size(builder.box, fit.width, fit.height, fit.align, r.fit_tr, r.fit_width, r.fit_height);
    info[i] = r;
}

__global__ void __launch_bounds__(kThreads) parse_emit_kernel(const uint8_t* __restrict__ text, const uint32_t* __restrict__ chunk_off,
                                                              uint32_t n_chunks, const ParseEmitBase* __restrict__ bases,
                                                              double* __restrict__ out_pts, uint2* __restrict__ out_items,
                                                              uint2* __restrict__ out_packed) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_chunks) return;
    const ParseEmitBase base = bases[i];
    if (base.pt == kParseSkip) return;
    const uint32_t a = chunk_off[i], b = chunk_off[i + 1];
    EmitOut<uint2> out;
    out.pts = out_pts;
    out.items = out_items;
    out.packed = out_packed;
    out.pt = base.pt;
    out.item = base.item;
    out.curve = base.curve;
    out.rest = base.rest;
    out.sub_first_pt = out.pt;
    out.closing_flag = kItemClosing;
    out.closed_flag = kItemExplicitClosed;
    PathBuild<EmitOut<uint2>> builder(out);
    uint32_t err_pos = 0;
    parse_svg_path(text + a, b - a, builder, err_pos);
}

}  // namespace

void launch_parse_count(const uint8_t* text, const uint32_t* chunk_off, uint32_t n_chunks, const ParseFit& fit, ParseInfoDev* info,
                        cudaStream_t s) {
    parse_count_kernel<<<(n_chunks + kThreads - 1) / kThreads, kThreads, 0, s>>>(text, chunk_off, n_chunks, fit, info);
}

void launch_parse_emit(const uint8_t* text, const uint32_t* chunk_off, uint32_t n_chunks, const ParseEmitBase* bases, double2* out_pts,
                       uint2* out_items, uint2* out_packed, cudaStream_t s) {
    parse_emit_kernel<<<(n_chunks + kThreads - 1) / kThreads, kThreads, 0, s>>>(text, chunk_off, n_chunks, bases,
                                                                                reinterpret_cast<double*>(out_pts), out_items, out_packed);
}

void parse_plan_chunks_host(const char* text, const uint32_t* text_off, uint32_t n_paths, std::vector<uint32_t>& chunk_off,
                            std::vector<uint32_t>& chunk_first) {
    parse_plan_chunks(text, text_off, n_paths, chunk_off, chunk_first);
}
void parse_merge_chunks_host(const ParseInfoDev* info, const std::vector<uint32_t>& chunk_off, const std::vector<uint32_t>& chunk_first,
                             const uint32_t* text_off, uint32_t n_paths, const ParseFit& fit, ParseInfoDev* path_info,
                             std::vector<ParseEmitBase>& bases, std::vector<uint32_t>& item_off, uint32_t& total_pts) {
    parse_merge_chunks(info, chunk_off, chunk_first, text_off, n_paths, fit, path_info, bases, item_off, total_pts);
}

}  // namespace rgpu
