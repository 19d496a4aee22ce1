// Tables of the batch SVG parser (parse.cu): pure C++, shared by the kernels, the library's host side and the test harness.
#pragma once
#include <cstdint>

namespace rgpu {

struct ParseFit {
    uint32_t width, height;
    int32_t align;  // < 0: no fit_size
};
struct ParseInfoDev {  // = rgpu_parse_info of the public header, plus n_curves in its padding word
    double bbox[4];
    double fit_tr[6];
    uint32_t fit_width, fit_height;
    uint32_t n_points, n_segments, n_subpaths;
    int32_t status;
    uint32_t error_offset;
    int32_t has_bbox;
    uint32_t n_curves, pad_;
};
struct ParseEmitBase {  // where a chunk writes: first point, first item (reference order), first curve / first other item (curves-first order)
    uint32_t pt, item, curve, rest;
};
constexpr uint32_t kParseSkip = 0xffffffffu;  // ParseEmitBase::pt of a chunk that emits nothing

}  // namespace rgpu
