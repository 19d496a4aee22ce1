// Host side of rgpu_path_stroke / rgpu_dpath_info / rgpu_dpath_download (included at the end of context.cu).

namespace {

int check_stroke_style(rgpu_ctx* ctx, const rgpu_stroke_style* style) {
    if (!style) return fail(ctx, RGPU_ERR_INVALID, "style is NULL");
    if (style->line_join < RGPU_JOIN_MITER || style->line_join > RGPU_JOIN_ROUND) return fail(ctx, RGPU_ERR_INVALID, "bad line_join");
    if (style->line_cap < RGPU_CAP_BUTT || style->line_cap > RGPU_CAP_ROUND) return fail(ctx, RGPU_ERR_INVALID, "bad line_cap");
    return RGPU_OK;
}

// The three passes over a unit table.  The source control points come from the host (`host_pts`, copied into the scratch
// block) or are already on the device (`dev_pts`; unit offsets index it directly).
int stroke_units_to_path(rgpu_ctx* ctx, const std::vector<StrokeUnit>& units, const double* host_pts, const double2* dev_pts, uint32_t n_points,
                         const rgpu_stroke_style* style, rgpu_dpath** out);

}  // namespace

int rgpu_path_stroke(rgpu_ctx* ctx, const rgpu_path* path, const rgpu_stroke_style* style, rgpu_dpath** out) {
    if (!ctx || !out) return RGPU_ERR_INVALID;
    *out = nullptr;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = validate_path(ctx, path);
    if (rc) return rc;
    if ((rc = check_stroke_style(ctx, style))) return rc;
    std::vector<StrokeUnit> units;
    build_stroke_units(path->kinds, path->n_segments, path->subpath_offsets, path->closed, path->n_subpaths, units);
    return stroke_units_to_path(ctx, units, path->points, nullptr, path->n_points, style, out);
}

int rgpu_dpath_stroke(rgpu_ctx* ctx, const rgpu_dpath* src, const rgpu_stroke_style* style, rgpu_dpath** out) {
    if (!ctx || !out || !src) return RGPU_ERR_INVALID;
    *out = nullptr;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = check_stroke_style(ctx, style))) return rc;
    // the path's structure comes from its item list (8 B per segment; the control points stay where they are)
    std::vector<uint2> items(src->n_items);
    if (src->n_items) CK(ctx, cudaMemcpyAsync(items.data(), src->items, sizeof(uint2) * src->n_items, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<StrokeUnit> units;
    units.reserve(2 * items.size());
    size_t first = 0;  // first item of the current subpath
    for (size_t i = 0; i < items.size(); i++) {
        if (!(items[i].y & kItemClosing)) continue;
        const bool closed = (items[i].y & kItemExplicitClosed) != 0;
        const uint32_t sp_start = items[i].y & kItemIndexMask, sp_end = items[i].x;
        uint32_t u0 = (uint32_t)units.size();
        for (size_t k = first; k < i; k++) units.push_back(StrokeUnit{items[k].x, items[k].y, u0, 0u});
        if (closed) {
            units.push_back(StrokeUnit{sp_start, kUnitCloser | (kCloserForward << 16), u0, sp_end});
            u0 = (uint32_t)units.size();
        }
        for (size_t k = i; k > first; k--) {
            const uint32_t cap = (!closed && k == i) ? kUnitCap : 0u;
            units.push_back(StrokeUnit{items[k - 1].x, items[k - 1].y | kUnitReversed | cap, u0, 0u});
        }
        units.push_back(StrokeUnit{sp_start, kUnitCloser | ((closed ? kCloserBackward : kCloserOpen) << 16), u0, sp_end});
        first = i + 1;
    }
    return stroke_units_to_path(ctx, units, nullptr, src->pts, src->n_points, style, out);
}

namespace {

int stroke_units_to_path(rgpu_ctx* ctx, const std::vector<StrokeUnit>& units, const double* host_pts, const double2* dev_pts, uint32_t n_points,
                         const rgpu_stroke_style* style, rgpu_dpath** out) {
    int rc;
    if (units.size() > 0x7fffffffull) return fail(ctx, RGPU_ERR_INVALID, "path too large");
    const uint32_t n = (uint32_t)units.size();
    auto* dp = new rgpu_dpath();
    if (n == 0) {
        *out = dp;
        return RGPU_OK;
    }
    // one scratch block: [units | source points | counts 4 x (n+1) | offsets 4 x (n+1) | first pieces | last pieces | scan temp]
    const size_t stride = stroke_count_stride(n);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_units = 0;
    const size_t o_pts = up(o_units + sizeof(StrokeUnit) * n);
    const size_t o_cnt = up(o_pts + (host_pts ? sizeof(double2) * n_points : 0));
    const size_t o_off = up(o_cnt + sizeof(uint32_t) * 4 * stride);
    const size_t o_first = up(o_off + sizeof(uint32_t) * 4 * stride);
    const size_t o_last = up(o_first + stroke_piece_bytes() * n);
    const size_t o_scan = up(o_last + stroke_piece_bytes() * n);
    const size_t scan_bytes = scan_temp_bytes(n + 1);
    const size_t total = o_scan + up(scan_bytes);
    auto bail = [&](int code) {
        free_path(dp);
        delete dp;
        return code;
    };
    if ((rc = ensure_dev(ctx, ctx->stroke_buf, total))) return bail(rc);
    char* base = static_cast<char*>(ctx->stroke_buf.p);
    auto* d_units = reinterpret_cast<StrokeUnit*>(base + o_units);
    const double2* d_pts = host_pts ? reinterpret_cast<const double2*>(base + o_pts) : dev_pts;
    auto* d_cnt = reinterpret_cast<uint32_t*>(base + o_cnt);
    auto* d_off = reinterpret_cast<uint32_t*>(base + o_off);
    cudaStream_t st = ctx->stream;
#define CKB(call)                                                            \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) {                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
            return bail(RGPU_ERR_CUDA);                                      \
        }                                                                    \
    } while (0)
    CKB(cudaMemcpyAsync(d_units, units.data(), sizeof(StrokeUnit) * n, cudaMemcpyHostToDevice, st));
    if (host_pts && n_points) CKB(cudaMemcpyAsync(base + o_pts, host_pts, sizeof(double2) * n_points, cudaMemcpyHostToDevice, st));
    CKB(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t) * 4 * stride, st));
    StrokeStyleDev sd{style->width, style->miter_limit, style->line_join, style->line_cap};
    // rgpu_set_profiling: stage 0 = pieces + count + scans, stage 1 = the host's turn (sizes, allocation), stage 2 = emit
    const bool prof = ctx->profiling && ctx->ev[0];
    ctx->ev_valid = false;
    if (prof) CKB(cudaEventRecord(ctx->ev[0], st));
    launch_stroke_pieces(d_units, n, d_pts, sd, d_cnt, base + o_first, base + o_last, st);
    launch_stroke_units(false, d_units, n, d_pts, sd, d_cnt, base + o_first, base + o_last, nullptr, nullptr, nullptr, nullptr, st);
    for (int k = 0; k < 4; k++) {
        CKB(cudaMemsetAsync(base + o_scan, 0, scan_bytes, st));
        launch_exclusive_scan(d_cnt + k * stride, d_off + k * stride, n + 1, base + o_scan, scan_bytes, st, true);
    }
    ctx->n_launches += 6;
    if (prof) CKB(cudaEventRecord(ctx->ev[1], st));
    uint32_t totals[4];
    for (int k = 0; k < 4; k++) CKB(cudaMemcpyAsync(&totals[k], d_off + k * stride + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CKB(cudaStreamSynchronize(st));
    CKB(cudaGetLastError());
    const uint32_t n_seg = totals[0], n_pts = totals[1], n_curves = totals[2], n_sub = totals[3];
    if (n_pts > kItemIndexMask) {
        ctx->err = "stroked path too large";
        return bail(RGPU_ERR_INVALID);
    }
    dp->n_points = n_pts;
    dp->n_items = n_seg + n_sub;
    dp->n_curves = n_curves;
    dp->n_segments = n_seg;
    dp->n_subpaths = n_sub;
    if (dp->n_items) {
        CKB(cudaMallocAsync(reinterpret_cast<void**>(&dp->pts), sizeof(double2) * std::max<uint32_t>(n_pts, 1), st));  // freed with cudaFree (free_path)
        CKB(cudaMallocAsync(reinterpret_cast<void**>(&dp->items), sizeof(uint2) * 2 * dp->n_items, st));
        dp->items_packed = dp->items + dp->n_items;
        if (prof) CKB(cudaEventRecord(ctx->ev[2], st));
        launch_stroke_units(true, d_units, n, d_pts, sd, d_cnt, base + o_first, base + o_last, d_off, dp->pts, dp->items, dp->items_packed, st);
        ctx->n_launches += 1;
        if (prof) {
            CKB(cudaEventRecord(ctx->ev[3], st));
            ctx->ev_valid = true;
        }
        CKB(cudaStreamSynchronize(st));
        CKB(cudaGetLastError());
    }
#undef CKB
    *out = dp;
    return RGPU_OK;
}

}  // namespace

int rgpu_dpath_info(const rgpu_dpath* p, uint32_t* n_points, uint32_t* n_segments, uint32_t* n_subpaths) {
    if (!p) return RGPU_ERR_INVALID;
    if (n_points) *n_points = p->n_points;
    if (n_segments) *n_segments = p->n_segments;
    if (n_subpaths) *n_subpaths = p->n_subpaths;
    return RGPU_OK;
}

int rgpu_dpath_download(rgpu_ctx* ctx, const rgpu_dpath* p, double* points, uint8_t* kinds, uint32_t* subpath_offsets, uint8_t* closed) {
    if (!ctx || !p) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if ((p->n_points && !points) || (p->n_segments && !kinds) || !subpath_offsets || (p->n_subpaths && !closed))
        return fail(ctx, RGPU_ERR_INVALID, "output arrays are NULL");
    std::vector<uint2> items(p->n_items);
    if (p->n_points) CK(ctx, cudaMemcpyAsync(points, p->pts, sizeof(double2) * p->n_points, cudaMemcpyDeviceToHost, ctx->stream));
    if (p->n_items) CK(ctx, cudaMemcpyAsync(items.data(), p->items, sizeof(uint2) * p->n_items, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t seg = 0, sub = 0;
    subpath_offsets[0] = 0;
    for (const uint2& it : items) {
        if (it.y & kItemClosing) {
            if (sub >= p->n_subpaths) return fail(ctx, RGPU_ERR_INVALID, "device path is inconsistent");
            closed[sub] = (it.y & kItemExplicitClosed) ? 1 : 0;
            subpath_offsets[++sub] = seg;
        } else {
            if (seg >= p->n_segments) return fail(ctx, RGPU_ERR_INVALID, "device path is inconsistent");
            kinds[seg++] = (uint8_t)it.y;
        }
    }
    if (seg != p->n_segments || sub != p->n_subpaths) return fail(ctx, RGPU_ERR_INVALID, "device path is inconsistent");
    return RGPU_OK;
}
