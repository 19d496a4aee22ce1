// Host side of the batch SVG parser's plan (see parse.cu): cutting long strings into chunks at absolute movetos, adding the
// chunks of a path up, planning where every chunk writes.  Pure C++ over parse_device.cuh (included first, with SD_FN
// covering the host), so that the test harness runs the very same code.
#pragma once
#include <cstring>
#include <vector>

#include "parse_device.cuh"
#include "parse_tables.hpp"

namespace rgpu {

using namespace sv;

// Host: cut long strings at absolute movetos.  chunk_off gets n_chunks + 1 byte offsets (the chunks tile the text),
// chunk_first[i] = first chunk of path i (n_paths + 1 entries).
inline void parse_plan_chunks(const char* text, const uint32_t* text_off, uint32_t n_paths, std::vector<uint32_t>& chunk_off,
                       std::vector<uint32_t>& chunk_first) {
    constexpr uint32_t kSplitAbove = 4096;  // strings up to this length stay whole
    constexpr uint32_t kMinChunk = 256;     // a cut needs this much text before it
    chunk_off.clear();
    chunk_first.resize((size_t)n_paths + 1);
    for (uint32_t i = 0; i < n_paths; i++) {
        const uint32_t a = text_off[i], b = text_off[i + 1];
        chunk_first[i] = (uint32_t)chunk_off.size();
        chunk_off.push_back(a);
        if (b - a <= kSplitAbove) continue;
        uint32_t start = a;
        while (true) {
            if (b - start <= kMinChunk) break;
            const void* m = std::memchr(text + start + kMinChunk, 'M', b - start - kMinChunk);
            if (!m) break;
            start = (uint32_t)(static_cast<const char*>(m) - text);
            chunk_off.push_back(start);
        }
    }
    chunk_first[n_paths] = (uint32_t)chunk_off.size();
    chunk_off.push_back(n_paths ? text_off[n_paths] : 0u);
}

// Host: add the chunks of every path up (`info` = per chunk in, `path_info` = per path out) and plan where every chunk
// writes.  A path with a failing chunk is empty and reports the first failure (the one the reference's serial parse
// meets); the boxes of the chunks are united (see DESIGN.md for the one-ulp caveat of not folding them serially); the
// fit of a path that spans several chunks is computed here with the same code the device runs.
inline void parse_merge_chunks(const ParseInfoDev* info, const std::vector<uint32_t>& chunk_off, const std::vector<uint32_t>& chunk_first,
                        const uint32_t* text_off, uint32_t n_paths, const ParseFit& fit, ParseInfoDev* path_info,
                        std::vector<ParseEmitBase>& bases, std::vector<uint32_t>& item_off, uint32_t& total_pts) {
    const uint32_t n_chunks = (uint32_t)chunk_off.size() - 1;
    bases.assign(n_chunks, ParseEmitBase{kParseSkip, 0u, 0u, 0u});
    item_off.resize((size_t)n_paths + 1);
    uint32_t pt = 0, item = 0;
    for (uint32_t i = 0; i < n_paths; i++) {
        const uint32_t c0 = chunk_first[i], c1 = chunk_first[i + 1];
        item_off[i] = item;
        ParseInfoDev r = info[c0];
        if (c1 - c0 > 1) {
            bool have = false;
            Box box;
            box.lo = box.hi = mk(0.0, 0.0);
            r.n_segments = r.n_subpaths = r.n_points = r.n_curves = 0;
            r.status = kParseOk;
            r.error_offset = 0;
            for (uint32_t c = c0; c < c1; c++) {
                const ParseInfoDev& q = info[c];
                if (q.status != kParseOk) {
                    r.status = q.status;
                    r.error_offset = q.error_offset + (chunk_off[c] - text_off[i]);
                    break;
                }
                r.n_segments += q.n_segments;
                r.n_subpaths += q.n_subpaths;
                r.n_points += q.n_points;
                r.n_curves += q.n_curves;
                if (q.has_bbox) {
                    const Box b = box_new(mk(q.bbox[0], q.bbox[1]), mk(q.bbox[2], q.bbox[3]));
                    box = have ? box_extend(box_extend(b, box.lo), box.hi) : b;
                    have = true;
                }
            }
            for (int k = 0; k < 6; k++) r.fit_tr[k] = (k == 0 || k == 4) ? 1.0 : 0.0;
            r.fit_width = r.fit_height = 0;
            if (r.status != kParseOk) {
                r.n_segments = r.n_subpaths = r.n_points = r.n_curves = 0;
                have = false;
            }
            r.has_bbox = have;
            r.bbox[0] = have ? box.lo.x : 0.0;
            r.bbox[1] = have ? box.lo.y : 0.0;
            r.bbox[2] = have ? box.hi.x : 0.0;
            r.bbox[3] = have ? box.hi.y : 0.0;
            if (have && fit.align >= 0) fit_size(box, fit.width, fit.height, fit.align, r.fit_tr, r.fit_width, r.fit_height);
        }
        path_info[i] = r;
        if (r.status == kParseOk && r.n_segments) {
            uint32_t curve = item, rest = item + r.n_curves;
            for (uint32_t c = c0; c < c1; c++) {
                const ParseInfoDev& q = info[c];
                if (!q.n_segments) continue;
                bases[c] = ParseEmitBase{pt, item, curve, rest};
                pt += q.n_points;
                item += q.n_segments + q.n_subpaths;
                curve += q.n_curves;
                rest += q.n_segments + q.n_subpaths - q.n_curves;
            }
        }
    }
    item_off[n_paths] = item;
    total_pts = pt;
}


}  // namespace rgpu
