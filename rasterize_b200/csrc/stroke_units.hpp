// Unit table of the device stroke (see stroke.cu): pure C++, shared by the library's host side and the kernels.
#pragma once
#include <cstdint>
#include <vector>

namespace rgpu {

// one thread per unit of the reference's serial walk (src/path.rs:374-415)
// ordinary unit: a = first control point of the source segment, b = kind | flags, c = first unit of its contour
// closer unit:   a = first point of the source subpath, d = its last point, b = kUnitCloser | mode << 16, c as above
struct StrokeUnit {
    uint32_t a, b, c, d;
};
constexpr uint32_t kUnitReversed = 1u << 8;  // the segment is walked backwards (`Curve::reverse`)
constexpr uint32_t kUnitCap = 1u << 9;       // `line_cap` instead of `line_join` in front of its pieces
constexpr uint32_t kUnitCloser = 1u << 15;
constexpr uint32_t kCloserForward = 0, kCloserBackward = 1, kCloserOpen = 2;

// The count / offset arrays of the passes (segments, points, curves, closed contours) hold n_units + 1 words each, padded
// so that every array starts on a 16-byte boundary.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t stroke_count_stride(uint32_t n_units) { return (n_units + 1u + 3u) & ~3u; }

// The reference's walk as a table of units; only the path's STRUCTURE is read (kinds, subpath offsets, closed flags).
inline void build_stroke_units(const uint8_t* kinds, uint32_t n_segments, const uint32_t* subpath_offsets, const uint8_t* closed,
                               uint32_t n_subpaths, std::vector<StrokeUnit>& units) {
    std::vector<uint32_t> pt_off((size_t)n_segments + 1);
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n_segments; i++) {
        pt_off[i] = acc;
        acc += kinds[i];
    }
    pt_off[n_segments] = acc;
    units.clear();
    units.reserve(2 * (size_t)n_segments + 2 * (size_t)n_subpaths);
    for (uint32_t s = 0; s < n_subpaths; s++) {
        const uint32_t a = subpath_offsets[s], b = subpath_offsets[s + 1];
        const uint32_t sp_start = pt_off[a], sp_end = pt_off[b] - 1;
        uint32_t u0 = (uint32_t)units.size();
        for (uint32_t i = a; i < b; i++) units.push_back(StrokeUnit{pt_off[i], kinds[i], u0, 0u});
        if (closed[s]) {
            units.push_back(StrokeUnit{sp_start, kUnitCloser | (kCloserForward << 16), u0, sp_end});
            u0 = (uint32_t)units.size();
        }
        for (uint32_t i = b; i > a; i--) {
            const uint32_t cap = (!closed[s] && i == b) ? kUnitCap : 0u;  // the turning point of an open subpath
            units.push_back(StrokeUnit{pt_off[i - 1], (uint32_t)kinds[i - 1] | kUnitReversed | cap, u0, 0u});
        }
        units.push_back(StrokeUnit{sp_start, kUnitCloser | ((closed[s] ? kCloserBackward : kCloserOpen) << 16), u0, sp_end});
    }
}

}  // namespace rgpu
