// Host side of rgpu_parse_svg_batch / rgpu_path_batch_info / rgpu_path_batch_download (included at the end of context.cu).

static_assert(sizeof(rgpu_parse_info) == sizeof(ParseInfoDev), "rgpu_parse_info and ParseInfoDev must have the same layout");

int rgpu_parse_svg_batch(rgpu_ctx* ctx, const char* text, const uint32_t* text_offsets, size_t n_paths, const rgpu_parse_options* opt,
                         rgpu_dpath_batch** out, rgpu_parse_info* info) {
    if (!ctx || !out) return RGPU_ERR_INVALID;
    *out = nullptr;
    CK(ctx, cudaSetDevice(ctx->device));
    if (n_paths && (!text_offsets || (!text && text_offsets[n_paths] != 0))) return fail(ctx, RGPU_ERR_INVALID, "text arrays are NULL");
    if (n_paths > 0x7fffffffull) return fail(ctx, RGPU_ERR_INVALID, "too many paths");
    for (size_t i = 0; i < n_paths; i++)
        if (text_offsets[i + 1] < text_offsets[i]) return fail(ctx, RGPU_ERR_INVALID, "text_offsets must not decrease");
    if (n_paths && text_offsets[0] != 0) return fail(ctx, RGPU_ERR_INVALID, "text_offsets must start at 0");
    ParseFit fit{0u, 0u, -1};
    if (opt && opt->fit_align >= 0) {
        if (opt->fit_align > RGPU_ALIGN_MAX) return fail(ctx, RGPU_ERR_INVALID, "bad fit_align");
        fit = ParseFit{opt->fit_width, opt->fit_height, opt->fit_align};
    }
    auto* b = new rgpu_dpath_batch();
    const uint32_t n = (uint32_t)n_paths;
    if (n == 0) {
        *out = b;
        return RGPU_OK;
    }
    const size_t text_bytes = text_offsets[n];
    std::vector<uint32_t> chunk_off, chunk_first;
    parse_plan_chunks_host(text, text_offsets, n, chunk_off, chunk_first);
    const uint32_t n_chunks = (uint32_t)chunk_off.size() - 1;
    // scratch: [text | chunk offsets | chunk infos | emit bases]
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_text = 0;
    const size_t o_coff = up(o_text + std::max<size_t>(text_bytes, 1));
    const size_t o_info = up(o_coff + sizeof(uint32_t) * ((size_t)n_chunks + 1));
    const size_t o_base = up(o_info + sizeof(ParseInfoDev) * n_chunks);
    const size_t total = o_base + up(sizeof(ParseEmitBase) * n_chunks);
    auto bail = [&](int code) {
        if (b->pts) cudaFree(b->pts);
        if (b->items) cudaFree(b->items);
        delete b;
        return code;
    };
    int rc;
    if ((rc = ensure_dev(ctx, ctx->stroke_buf, total))) return bail(rc);
    char* base = static_cast<char*>(ctx->stroke_buf.p);
    auto* d_text = reinterpret_cast<uint8_t*>(base + o_text);
    auto* d_coff = reinterpret_cast<uint32_t*>(base + o_coff);
    auto* d_info = reinterpret_cast<ParseInfoDev*>(base + o_info);
    auto* d_base = reinterpret_cast<ParseEmitBase*>(base + o_base);
    cudaStream_t st = ctx->stream;
#define CKB(call)                                                            \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) {                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
            return bail(RGPU_ERR_CUDA);                                      \
        }                                                                    \
    } while (0)
    if (text_bytes) CKB(cudaMemcpyAsync(d_text, text, text_bytes, cudaMemcpyHostToDevice, st));
    CKB(cudaMemcpyAsync(d_coff, chunk_off.data(), sizeof(uint32_t) * ((size_t)n_chunks + 1), cudaMemcpyHostToDevice, st));
    // rgpu_set_profiling: stage 0 = count, stage 1 = the host's turn (adding up the chunks, allocation), stage 2 = emit
    const bool prof = ctx->profiling && ctx->ev[0];
    ctx->ev_valid = false;
    if (prof) CKB(cudaEventRecord(ctx->ev[0], st));
    launch_parse_count(d_text, d_coff, n_chunks, fit, d_info, st);
    ctx->n_launches += 1;
    if (prof) CKB(cudaEventRecord(ctx->ev[1], st));
    // per-chunk results -> per-path results: in the caller's table, or a temporary one (the views need item ranges and curve counts)
    std::vector<ParseInfoDev> chunk_info_store;
    std::vector<rgpu_parse_info> local;
    rgpu_parse_info* h_info = info;
    if (!h_info) {
        local.resize(n);
        h_info = local.data();
    }
    ParseInfoDev* h_path = reinterpret_cast<ParseInfoDev*>(h_info);
    ParseInfoDev* h_chunk = h_path;  // no string was cut: the chunk table is the path table
    if (n_chunks != n) {
        chunk_info_store.resize(n_chunks);
        h_chunk = chunk_info_store.data();
    }
    CKB(cudaMemcpyAsync(h_chunk, d_info, sizeof(ParseInfoDev) * n_chunks, cudaMemcpyDeviceToHost, st));
    CKB(cudaStreamSynchronize(st));
    CKB(cudaGetLastError());
    std::vector<ParseEmitBase> bases;
    std::vector<uint32_t> item_off;
    uint32_t total_pts = 0;
    parse_merge_chunks_host(h_chunk, chunk_off, chunk_first, text_offsets, n, fit, h_path, bases, item_off, total_pts);
    const uint32_t n_items = item_off[n];
    if (total_pts > kItemIndexMask) {
        ctx->err = "parsed batch too large";
        return bail(RGPU_ERR_INVALID);
    }
    if (!info) {
        for (uint32_t i = 0; i < n; i++)
            if (h_info[i].status) {
                static const char* kinds[] = {"", "InvalidCmd", "InvalidScalar", "InvalidFlag"};
                ctx->err = "path " + std::to_string(i) + ": " + kinds[h_info[i].status & 3] + " at offset " + std::to_string(h_info[i].error_offset);
                return bail(RGPU_ERR_INVALID);
            }
    }
    if (n_items) {
        CKB(cudaMallocAsync(reinterpret_cast<void**>(&b->pts), sizeof(double2) * std::max<uint32_t>(total_pts, 1), st));  // freed with cudaFree
        CKB(cudaMallocAsync(reinterpret_cast<void**>(&b->items), sizeof(uint2) * 2 * n_items, st));
        CKB(cudaMemcpyAsync(d_base, bases.data(), sizeof(ParseEmitBase) * n_chunks, cudaMemcpyHostToDevice, st));
        if (prof) CKB(cudaEventRecord(ctx->ev[2], st));
        launch_parse_emit(d_text, d_coff, n_chunks, d_base, b->pts, b->items, b->items + n_items, st);
        ctx->n_launches += 1;
        if (prof) {
            CKB(cudaEventRecord(ctx->ev[3], st));
            ctx->ev_valid = true;
        }
    }
    b->n_points = total_pts;
    b->n_items = n_items;
    b->paths.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        rgpu_dpath& d = b->paths[i];
        d.pts = b->pts;
        d.items = b->items ? b->items + item_off[i] : nullptr;
        d.items_packed = b->items ? b->items + n_items + item_off[i] : nullptr;
        d.n_items = item_off[i + 1] - item_off[i];
        d.n_curves = h_path[i].n_curves;
        d.n_points = total_pts;
        b->n_segments += h_info[i].n_segments;
        b->n_subpaths += h_info[i].n_subpaths;
    }
    CKB(cudaStreamSynchronize(st));  // `bases` and the chunk table die with this call
    CKB(cudaGetLastError());
#undef CKB
    *out = b;
    return RGPU_OK;
}

int rgpu_path_batch_info(const rgpu_dpath_batch* b, size_t* n_paths, uint32_t* n_points, uint32_t* n_segments, uint32_t* n_subpaths) {
    if (!b) return RGPU_ERR_INVALID;
    if (n_paths) *n_paths = b->paths.size();
    if (n_points) *n_points = b->n_points;
    if (n_segments) *n_segments = b->n_segments;
    if (n_subpaths) *n_subpaths = b->n_subpaths;
    return RGPU_OK;
}

int rgpu_path_batch_download(rgpu_ctx* ctx, const rgpu_dpath_batch* b, double* points, uint8_t* kinds, uint32_t* subpath_offsets, uint8_t* closed,
                             uint32_t* path_subpath_offsets) {
    if (!ctx || !b) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if ((b->n_points && !points) || (b->n_segments && !kinds) || !subpath_offsets || (b->n_subpaths && !closed) || !path_subpath_offsets)
        return fail(ctx, RGPU_ERR_INVALID, "output arrays are NULL");
    std::vector<uint2> items(b->n_items);
    if (b->n_points) CK(ctx, cudaMemcpyAsync(points, b->pts, sizeof(double2) * b->n_points, cudaMemcpyDeviceToHost, ctx->stream));
    if (b->n_items) CK(ctx, cudaMemcpyAsync(items.data(), b->items, sizeof(uint2) * b->n_items, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t seg = 0, sub = 0;
    subpath_offsets[0] = 0;
    path_subpath_offsets[0] = 0;
    for (size_t i = 0; i < b->paths.size(); i++) {
        const rgpu_dpath& d = b->paths[i];
        const size_t first = d.items ? (size_t)(d.items - b->items) : 0;
        for (size_t k = first; k < first + d.n_items; k++) {
            const uint2 it = items[k];
            if (it.y & kItemClosing) {
                if (sub >= b->n_subpaths) return fail(ctx, RGPU_ERR_INVALID, "device batch is inconsistent");
                closed[sub] = (it.y & kItemExplicitClosed) ? 1 : 0;
                subpath_offsets[++sub] = seg;
            } else {
                if (seg >= b->n_segments) return fail(ctx, RGPU_ERR_INVALID, "device batch is inconsistent");
                kinds[seg++] = (uint8_t)it.y;
            }
        }
        path_subpath_offsets[i + 1] = sub;
    }
    if (seg != b->n_segments || sub != b->n_subpaths) return fail(ctx, RGPU_ERR_INVALID, "device batch is inconsistent");
    return RGPU_OK;
}
