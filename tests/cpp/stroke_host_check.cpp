// TEST HARNESS: the per-unit stroke code of the device path (rasterize_b200/csrc/stroke_device.cuh, the very functions the
// kernels of stroke.cu call) compiled for the host and run one unit after the other, so the decomposition of the
// reference's serial walk into units — tables, look-behind, counts, offsets, order of the output — can be checked
// against the oracle without a GPU.  Not part of the library; built by tests/test_stroke_units.py with
//   g++ -O1 -ffp-contract=off -shared -fPIC
#define SD_FN
#include <cstring>
#include <vector>

#include "../../rasterize_b200/csrc/stroke_device.cuh"

using namespace rgpu;
using namespace rgpu::sk;

namespace {
struct I2 {
    uint32_t x, y;
};
std::vector<double> g_pts;
std::vector<uint8_t> g_kinds, g_closed;
std::vector<uint32_t> g_sp;
}  // namespace

extern "C" int stroke_check_run(const double* points, const uint8_t* kinds, uint32_t n_segments, const uint32_t* sp_off, const uint8_t* closed,
                                uint32_t n_subpaths, double width, double miter_limit, int join, int cap, uint32_t* n_pts_out,
                                uint32_t* n_seg_out, uint32_t* n_sub_out) {
    std::vector<StrokeUnit> units;
    build_stroke_units(kinds, n_segments, sp_off, closed, n_subpaths, units);
    const uint32_t n = (uint32_t)units.size();
    Style style{width, miter_limit, join, cap};
    const size_t stride = stroke_count_stride(n);
    std::vector<uint32_t> cnt(4 * stride, 0u), off(4 * stride, 0u);
    std::vector<PieceRec> first(n), last(n);
    for (uint32_t i = 0; i < n; i++)
        unit_pieces(i, units.data(), points, style, cnt.data(), cnt.data() + stride, cnt.data() + 2 * stride, first.data(), last.data());
    // the count pass in REVERSE order: its result must not depend on which units have already added their join
    for (uint32_t i = n; i-- > 0;)
        unit_count(i, units.data(), n, points, style, cnt.data(), cnt.data() + stride, cnt.data() + 2 * stride, cnt.data() + 3 * stride,
                   first.data(), last.data());
    for (int k = 0; k < 4; k++) {
        uint32_t acc = 0;
        for (size_t i = 0; i <= n; i++) {
            off[k * stride + i] = acc;
            acc += cnt[k * stride + i];
        }
    }
    const uint32_t n_seg = off[n], n_pts = off[stride + n], n_curves = off[2 * stride + n], n_sub = off[3 * stride + n];
    const uint32_t n_items = n_seg + n_sub;
    g_pts.assign(2 * (size_t)n_pts, 0.0);
    std::vector<I2> items(n_items, I2{0xffffffffu, 0u}), packed(n_items, I2{0xffffffffu, 0u});
    for (uint32_t i = 0; i < n; i++) {
        EmitSink<I2> sink;
        sink.pts = g_pts.data();
        sink.items = items.data();
        sink.packed = packed.data();
        sink.pt = off[stride + i];
        sink.item = off[i] + off[3 * stride + i];
        sink.curve = off[2 * stride + i];
        sink.total_curves = n_curves;
        unit_emit(i, units.data(), n, points, style, cnt.data(), first.data(), last.data(), off[stride + units[i].c], 0xC0000000u, sink);
        // every unit must end exactly where the next one starts
        if (sink.pt != off[stride + i + 1] || sink.item != off[i + 1] + off[3 * stride + i + 1] || sink.curve != off[2 * stride + i + 1]) return -2;
    }
    // the curves-first list must be a permutation of the item list: curves in order, then the rest in order
    {
        size_t c = 0, r = n_curves;
        for (const I2& it : items) {
            if (it.x == 0xffffffffu) return -3;
            const bool curve = !(it.y & 0x80000000u) && it.y != 2u;
            const I2& q = packed[curve ? c++ : r++];
            if (q.x != it.x || q.y != it.y) return -4;
        }
        if (c != n_curves || r != n_items) return -5;
    }
    g_kinds.clear();
    g_closed.clear();
    g_sp.assign(1, 0u);
    uint32_t start_pt = 0;
    for (const I2& it : items) {
        if (it.y & 0x80000000u) {
            if ((it.y & 0x3fffffffu) != start_pt) return -6;  // closing item: back to the contour's first point
            g_closed.push_back((it.y & 0x40000000u) ? 1 : 0);
            g_sp.push_back((uint32_t)g_kinds.size());
            start_pt = it.x + 1;
        } else {
            g_kinds.push_back((uint8_t)it.y);
        }
    }
    if (g_kinds.size() != n_seg || g_closed.size() != n_sub) return -7;
    *n_pts_out = n_pts;
    *n_seg_out = n_seg;
    *n_sub_out = n_sub;
    return 0;
}

extern "C" void stroke_check_fetch(double* points, uint8_t* kinds, uint32_t* sp_off, uint8_t* closed) {
    if (!g_pts.empty()) std::memcpy(points, g_pts.data(), sizeof(double) * g_pts.size());
    if (!g_kinds.empty()) std::memcpy(kinds, g_kinds.data(), g_kinds.size());
    std::memcpy(sp_off, g_sp.data(), sizeof(uint32_t) * g_sp.size());
    if (!g_closed.empty()) std::memcpy(closed, g_closed.data(), g_closed.size());
}

extern "C" double stroke_check_hypot(double x, double y) { return hypot_libm(x, y); }
