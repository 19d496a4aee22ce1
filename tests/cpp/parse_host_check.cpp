// TEST HARNESS: the device SVG path parser (rasterize_b200/csrc/parse_device.cuh — the very code the kernels of parse.cu
// call) compiled for the host and run path by path through the same two passes (count, scan, emit), so the parser, the
// builder semantics, the bbox / fit_size arithmetic and the layout of the emitted batch can be checked against the oracle
// without a GPU.  Not part of the library; built by tests/test_parse_units.py with g++ -O1 -ffp-contract=off -shared -fPIC.
#define SD_FN
#include <cstring>
#include <vector>

#include "../../rasterize_b200/csrc/parse_device.cuh"
#include "../../rasterize_b200/csrc/parse_plan.hpp"

using namespace rgpu;
using namespace rgpu::sv;

namespace {
struct I2 {
    uint32_t x, y;
};
std::vector<double> g_pts;
std::vector<uint8_t> g_kinds, g_closed;
std::vector<uint32_t> g_sp, g_psp;
}  // namespace

// info_out: n records of rgpu_parse_info (= ParseInfoDev); returns 0 or a negative self-check code; *n_chunks_out = chunks used
extern "C" int parse_check_run(const char* text, const uint32_t* text_off, uint32_t n, uint32_t fit_w, uint32_t fit_h, int fit_align, void* info_out,
                               uint32_t* n_pts_out, uint32_t* n_seg_out, uint32_t* n_sub_out, uint32_t* n_chunks_out) {
    ParseInfoDev* path_info = static_cast<ParseInfoDev*>(info_out);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(text);
    const ParseFit fit{fit_w, fit_h, fit_align};
    std::vector<uint32_t> chunk_off, chunk_first;
    parse_plan_chunks(text, text_off, n, chunk_off, chunk_first);
    const uint32_t n_chunks = (uint32_t)chunk_off.size() - 1;
    *n_chunks_out = n_chunks;
    std::vector<ParseInfoDev> info(n_chunks);
    for (uint32_t i = 0; i < n_chunks; i++) {  // = parse_count_kernel
        CountOut out;
        PathBuild<CountOut> builder(out);
        uint32_t err_pos = 0;
        const int status = parse_svg_path(bytes + chunk_off[i], chunk_off[i + 1] - chunk_off[i], builder, err_pos);
        ParseInfoDev& r = info[i];
        std::memset(&r, 0, sizeof(r));
        r.status = status;
        r.error_offset = status ? err_pos : 0u;
        const bool ok = status == kParseOk;
        r.n_segments = ok ? out.seg : 0u;
        r.n_subpaths = ok ? out.sub : 0u;
        r.n_points = ok ? out.pts : 0u;
        r.n_curves = ok ? out.curves : 0u;
        r.has_bbox = ok && builder.has_box;
        if (r.has_bbox) {
            r.bbox[0] = builder.box.lo.x;
            r.bbox[1] = builder.box.lo.y;
            r.bbox[2] = builder.box.hi.x;
            r.bbox[3] = builder.box.hi.y;
        }
        r.fit_tr[0] = r.fit_tr[4] = 1.0;
        if (r.has_bbox && fit_align >= 0) fit_size(builder.box, fit_w, fit_h, fit_align, r.fit_tr, r.fit_width, r.fit_height);
    }
    std::vector<ParseEmitBase> bases;
    std::vector<uint32_t> off_items;
    uint32_t n_pts = 0;
    std::vector<ParseInfoDev> merged(n);
    parse_merge_chunks(info.data(), chunk_off, chunk_first, text_off, n, fit, merged.data(), bases, off_items, n_pts);
    for (uint32_t i = 0; i < n; i++) path_info[i] = merged[i];
    const uint32_t n_items = n ? off_items[n] : 0;
    g_pts.assign(2 * (size_t)n_pts, 0.0);
    std::vector<I2> items(n_items, I2{0xffffffffu, 0u}), packed(n_items, I2{0xffffffffu, 0u});
    for (uint32_t i = 0; i < n_chunks; i++) {  // = parse_emit_kernel
        if (bases[i].pt == kParseSkip) continue;
        EmitOut<I2> out;
        out.pts = g_pts.data();
        out.items = items.data();
        out.packed = packed.data();
        out.pt = bases[i].pt;
        out.item = bases[i].item;
        out.curve = bases[i].curve;
        out.rest = bases[i].rest;
        out.sub_first_pt = out.pt;
        out.closing_flag = 0x80000000u;
        out.closed_flag = 0x40000000u;
        PathBuild<EmitOut<I2>> builder(out);
        uint32_t err_pos = 0;
        parse_svg_path(bytes + chunk_off[i], chunk_off[i + 1] - chunk_off[i], builder, err_pos);
        if (out.pt != bases[i].pt + info[i].n_points || out.item != bases[i].item + info[i].n_segments + info[i].n_subpaths ||
            out.curve != bases[i].curve + info[i].n_curves)
            return -2;  // the two passes disagree
    }
    g_kinds.clear();
    g_closed.clear();
    g_sp.assign(1, 0u);
    g_psp.assign(1, 0u);
    for (uint32_t i = 0; i < n; i++) {
        // the curves-first list of a path: its curves in order, then the rest in order
        size_t c = off_items[i], r = off_items[i] + path_info[i].n_curves;
        for (uint32_t k = off_items[i]; k < off_items[i + 1]; k++) {
            const I2 it = items[k];
            if (it.x == 0xffffffffu) return -3;
            const bool closing = (it.y & 0x80000000u) != 0;
            const I2& q = packed[(!closing && it.y != 2u) ? c++ : r++];
            if (q.x != it.x || q.y != it.y) return -4;
            if (closing) {
                g_closed.push_back((it.y & 0x40000000u) ? 1 : 0);
                g_sp.push_back((uint32_t)g_kinds.size());
            } else {
                g_kinds.push_back((uint8_t)it.y);
            }
        }
        g_psp.push_back((uint32_t)g_closed.size());
    }
    *n_pts_out = n_pts;
    *n_seg_out = (uint32_t)g_kinds.size();
    *n_sub_out = (uint32_t)g_closed.size();
    return 0;
}

extern "C" void parse_check_fetch(double* points, uint8_t* kinds, uint32_t* sp_off, uint8_t* closed, uint32_t* path_sp_off) {
    if (!g_pts.empty()) std::memcpy(points, g_pts.data(), sizeof(double) * g_pts.size());
    if (!g_kinds.empty()) std::memcpy(kinds, g_kinds.data(), g_kinds.size());
    std::memcpy(sp_off, g_sp.data(), sizeof(uint32_t) * g_sp.size());
    if (!g_closed.empty()) std::memcpy(closed, g_closed.data(), g_closed.size());
    std::memcpy(path_sp_off, g_psp.data(), sizeof(uint32_t) * g_psp.size());
}
