// Batches of independent paths, prepared batches, band-sharded masks and the several-GPU front end of the C ABI
// (include/rasterize_b200.h, "batches of independent paths" / "several GPUs of one box").  Included at the end of
// context.cu: it uses that file's context structure and helpers.  SURVEY §8e: both sharding modes are exchange-free —
// a batch shards by path, a huge canvas by bands of rows (rows are independent in the signed-difference rasterizer,
// reference src/rasterize.rs:421-469, 478-503) — and every device returns its shard over its own PCIe link.

struct rgpu_dpath_batch {
    std::vector<rgpu_dpath> paths;  // views into the two allocations below
    double2* pts = nullptr;
    uint2* items = nullptr;         // [reference order of every path | curves-first order of every path]
    uint32_t n_points = 0, n_items = 0, n_segments = 0, n_subpaths = 0;  // totals (rgpu_path_batch_info / _download)
};

struct rgpu_batch {
    const rgpu_job* jobs = nullptr;  // caller-owned
    size_t n_jobs = 0;
    uint32_t flags = 0;
    bool small = false;              // every canvas fits the fused small-canvas kernel: tables live on the device
    JobDev* d_jobs = nullptr;
    PaintDev* d_paints = nullptr;
    uint32_t n_live = 0;
    bool gradients = false;
};

namespace {

int validate_batch(rgpu_ctx* ctx, const rgpu_path* all, const uint32_t* pso, size_t n_paths) {
    int rc = validate_path(ctx, all);
    if (rc) return rc;
    if (n_paths && !pso) return fail(ctx, RGPU_ERR_INVALID, "path_subpath_offsets is NULL");
    if (n_paths > 0x7fffffffull) return fail(ctx, RGPU_ERR_INVALID, "too many paths");
    if (n_paths) {
        if (pso[0] != 0 || pso[n_paths] != all->n_subpaths) return fail(ctx, RGPU_ERR_INVALID, "path_subpath_offsets must start at 0 and end at n_subpaths");
        for (size_t i = 0; i < n_paths; i++)
            if (pso[i + 1] < pso[i]) return fail(ctx, RGPU_ERR_INVALID, "path_subpath_offsets must not decrease");
    }
    return RGPU_OK;
}

// first point of every segment (+ the total)
void segment_point_offsets(const rgpu_path* p, std::vector<uint32_t>& off) {
    off.resize((size_t)p->n_segments + 1);
    uint32_t acc = 0;
    for (uint32_t i = 0; i < p->n_segments; i++) {
        off[i] = acc;
        acc += p->kinds[i];
    }
    off[p->n_segments] = acc;
}

// Item lists of paths [a, b) of a flat batch, appended to `ref` (the reference's emission order, src/path.rs:761-795)
// and `packed` (curves first): path i owns [item_off[i - a], item_off[i - a + 1]) of both.  Point indices are relative
// to point `pt_base` (the first point of path a), so a chunk can be uploaded on its own.
void build_range_items(const rgpu_path* all, const uint32_t* pso, const std::vector<uint32_t>& pt_off, size_t a, size_t b, uint32_t pt_base,
                       std::vector<uint2>& ref, std::vector<uint2>& packed, std::vector<uint32_t>& item_off, std::vector<uint32_t>& n_curves) {
    item_off.assign(1, (uint32_t)ref.size());
    n_curves.clear();
    for (size_t i = a; i < b; i++) {
        const size_t first = ref.size();
        uint32_t nc = 0;
        for (uint32_t s = pso[i]; s < pso[i + 1]; s++) {
            const uint32_t sa = all->subpath_offsets[s], sb = all->subpath_offsets[s + 1];
            for (uint32_t k = sa; k < sb; k++) {
                ref.push_back(make_uint2(pt_off[k] - pt_base, all->kinds[k]));
                if (all->kinds[k] != 2) nc++;
            }
            ref.push_back(make_uint2(pt_off[sb] - 1 - pt_base, kItemClosing | (all->closed[s] ? kItemExplicitClosed : 0u) | (pt_off[sa] - pt_base)));
        }
        for (size_t k = first; k < ref.size(); k++)
            if (!(ref[k].y & kItemClosing) && ref[k].y != 2u) packed.push_back(ref[k]);
        for (size_t k = first; k < ref.size(); k++)
            if ((ref[k].y & kItemClosing) || ref[k].y == 2u) packed.push_back(ref[k]);
        n_curves.push_back(nc);
        item_off.push_back((uint32_t)ref.size());
    }
}

// Host side of the split download of rgpu_fill_batch_host: `n_px` pixels of coverage (f32) become premultiplied LinColor
// pixels colour * alpha — the very multiplication the kernel does for a plain solid paint (small.cu finish_rows), so the
// bytes are the ones the device would have sent.  Non-temporal stores: the destination is written once and not read here.
void expand_alpha(const float* alpha, const float colour[4], float* out, size_t n_px) { rgpu::expand_alpha_simd(alpha, colour, out, n_px); }

// the condition under which RGPU_JOB_RENDER writes colour * alpha (small.cu: `plain`)
bool plain_solid(const rgpu_paint* p) {
    if (!p || p->kind != RGPU_PAINT_SOLID) return false;
    for (int k = 0; k < 4; k++)
        if (!(p->solid[k] >= 0.0f) || std::signbit(p->solid[k]) || !(p->solid[k] < 3e38f)) return false;
    return true;
}

size_t out_elem_bytes(int fmt) { return fmt == RGPU_OUT_LINCOLOR ? 16 : 4; }

}  // namespace

extern "C" {

int rgpu_path_upload_batch(rgpu_ctx* ctx, const rgpu_path* all, const uint32_t* pso, size_t n_paths, rgpu_dpath_batch** out) {
    if (!ctx || !out) return RGPU_ERR_INVALID;
    *out = nullptr;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = validate_batch(ctx, all, pso, n_paths);
    if (rc) return rc;
    std::vector<uint32_t> pt_off, item_off, n_curves;
    segment_point_offsets(all, pt_off);
    std::vector<uint2> ref, packed;
    ref.reserve((size_t)all->n_segments + all->n_subpaths);
    packed.reserve((size_t)all->n_segments + all->n_subpaths);
    build_range_items(all, pso, pt_off, 0, n_paths, 0, ref, packed, item_off, n_curves);
    auto* b = new rgpu_dpath_batch();
    const size_t n_items = ref.size();
    cudaError_t e = cudaSuccess;
    if (all->n_points) e = cudaMalloc(reinterpret_cast<void**>(&b->pts), sizeof(double2) * all->n_points);
    if (e == cudaSuccess && n_items) e = cudaMalloc(reinterpret_cast<void**>(&b->items), sizeof(uint2) * 2 * n_items);
    if (e == cudaSuccess && all->n_points) e = cudaMemcpyAsync(b->pts, all->points, sizeof(double2) * all->n_points, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n_items) e = cudaMemcpyAsync(b->items, ref.data(), sizeof(uint2) * n_items, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n_items) e = cudaMemcpyAsync(b->items + n_items, packed.data(), sizeof(uint2) * n_items, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // the host vectors die with this call
    if (e != cudaSuccess) {
        ctx->err = std::string("rgpu_path_upload_batch: ") + cudaGetErrorString(e);
        if (b->pts) cudaFree(b->pts);
        if (b->items) cudaFree(b->items);
        delete b;
        return RGPU_ERR_CUDA;
    }
    b->n_points = all->n_points;
    b->n_items = (uint32_t)n_items;
    b->n_segments = all->n_segments;
    b->n_subpaths = all->n_subpaths;
    b->paths.resize(n_paths);
    for (size_t i = 0; i < n_paths; i++) {
        rgpu_dpath& d = b->paths[i];
        d.pts = b->pts;
        d.items = b->items + item_off[i];
        d.items_packed = b->items + n_items + item_off[i];
        d.n_items = item_off[i + 1] - item_off[i];
        d.n_curves = n_curves[i];
        d.n_points = all->n_points;
    }
    *out = b;
    return RGPU_OK;
}

const rgpu_dpath* rgpu_path_batch_get(const rgpu_dpath_batch* batch, size_t i) {
    return (batch && i < batch->paths.size()) ? &batch->paths[i] : nullptr;
}

void rgpu_path_batch_free(rgpu_ctx* ctx, rgpu_dpath_batch* b) {
    if (!b) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (b->pts) cudaFree(b->pts);
    if (b->items) cudaFree(b->items);
    delete b;
}

int rgpu_batch_create(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags, rgpu_batch** out) {
    if (!ctx || !out || (!jobs && n_jobs)) return RGPU_ERR_INVALID;
    *out = nullptr;
    CK(ctx, cudaSetDevice(ctx->device));
    auto* b = new rgpu_batch();
    b->jobs = jobs;
    b->n_jobs = n_jobs;
    b->flags = flags;
    if (n_jobs && (flags & RGPU_BATCH_INDEPENDENT)) {
        Tables tb;
        int rc = build_tables(ctx, jobs, n_jobs, 1, nullptr, tb);
        if (rc) {
            delete b;
            return rc;
        }
        if (tb.all_small && tb.n_live) {
            cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&b->d_jobs), sizeof(JobDev) * tb.n_live);
            if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->d_paints), sizeof(PaintDev) * std::max<uint32_t>(tb.n_paints, 1));
            if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_jobs, ctx->h_jobs, sizeof(JobDev) * tb.n_live, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess && tb.n_paints)
                e = cudaMemcpyAsync(b->d_paints, ctx->h_paints, sizeof(PaintDev) * tb.n_paints, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->h_tables_ev, ctx->stream);
            if (e != cudaSuccess) {
                ctx->err = std::string("rgpu_batch_create: ") + cudaGetErrorString(e);
                if (b->d_jobs) cudaFree(b->d_jobs);
                if (b->d_paints) cudaFree(b->d_paints);
                delete b;
                return RGPU_ERR_CUDA;
            }
            b->small = true;
            b->n_live = tb.n_live;
            b->gradients = tb.gradients;
        }
    }
    *out = b;
    return RGPU_OK;
}

int rgpu_batch_render(rgpu_ctx* ctx, rgpu_batch* b) {
    if (!ctx || !b) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (!b->small) return submit(ctx, b->jobs, b->n_jobs, b->flags, 1);
    if (!(ctx->flatness > 0.0)) return fail(ctx, RGPU_ERR_INVALID, "flatness must be > 0 (the reference loops forever on 0)");
    cudaStream_t s = ctx->stream;
    Status* d_status = static_cast<Status*>(ctx->status.p);
    CK(ctx, cudaMemsetAsync(d_status, 0, sizeof(Status), s));
    const bool prof = ctx->profiling;
    ctx->ev_valid = false;
    if (prof)
        for (int i = 0; i < 3; i++) CK(ctx, cudaEventRecord(ctx->ev[i], s));
    launch_small_canvas(b->d_jobs, 0, b->n_live, b->d_paints, 16.0 * ctx->flatness * ctx->flatness, d_status, b->gradients, s);
    ctx->n_launches += 1;
    if (prof) {
        CK(ctx, cudaEventRecord(ctx->ev[3], s));
        ctx->ev_valid = true;
    }
    ctx->need_lines = ctx->need_refs = 0;
    ctx->d_status_cur = d_status;
    CK(ctx, cudaGetLastError());
    return RGPU_OK;
}

void rgpu_batch_free(rgpu_ctx* ctx, rgpu_batch* b) {
    if (!b) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (b->d_jobs) cudaFree(b->d_jobs);
    if (b->d_paints) cudaFree(b->d_paints);
    delete b;
}

int rgpu_fill_batch_host(rgpu_ctx* ctx, const rgpu_path* all, const uint32_t* pso, size_t n_paths, const double* trs, int fill_rule,
                         const rgpu_paint* paint, uint32_t width, uint32_t height, int out_format, void* out_host) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = validate_batch(ctx, all, pso, n_paths);
    if (rc) return rc;
    if (out_format < RGPU_OUT_LINCOLOR || out_format > RGPU_OUT_COVERAGE) return fail(ctx, RGPU_ERR_INVALID, "unknown output format");
    if (fill_rule != RGPU_NONZERO && fill_rule != RGPU_EVENODD) return fail(ctx, RGPU_ERR_INVALID, "bad fill rule");
    const bool coverage = out_format == RGPU_OUT_COVERAGE;
    if (!coverage && !paint) return fail(ctx, RGPU_ERR_INVALID, "paint is NULL");
    if (!coverage && paint->kind != RGPU_PAINT_SOLID && paint->units == RGPU_UNITS_BOUNDING_BOX)
        return fail(ctx, RGPU_ERR_INVALID, "bounding-box paint units need one bbox per path: use rgpu_render_batch");
    ctx->last_h2d_bytes = ctx->last_d2h_bytes = 0;
    if (n_paths == 0 || width == 0 || height == 0) return RGPU_OK;
    if (!out_host) return fail(ctx, RGPU_ERR_INVALID, "out_host is NULL");
    if (!ctx->copy_stream) {
        CK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < rgpu_ctx::kRing; i++) {
            CK(ctx, cudaEventCreateWithFlags(&ctx->ring_done[i], cudaEventDisableTiming));
            CK(ctx, cudaEventCreateWithFlags(&ctx->ring_copied[i], cudaEventDisableTiming));
        }
    }
    const auto t_call = std::chrono::steady_clock::now();
    const size_t px = (size_t)width * height;
    const size_t slab_elem = coverage ? 4 : 16;                  // what the kernels write
    const size_t out_elem = out_elem_bytes(out_format);          // what crosses PCIe
    // chunks of about 128 MB of kernel output: large enough to hide launch and copy set-up, small enough that the first
    // download starts early and three slabs stay modest
    static const size_t chunk_mb = getenv("RGPU_E2E_CHUNK_MB") ? (size_t)std::max(1, atoi(getenv("RGPU_E2E_CHUNK_MB"))) : 128;  // tuning
    size_t chunk = std::max<size_t>(1, (chunk_mb << 20) / (px * slab_elem));
    chunk = std::min(chunk, n_paths);
    const size_t n_chunks = (n_paths + chunk - 1) / chunk;
    const int ring = (int)std::min<size_t>(rgpu_ctx::kRing, n_chunks);
    // Split download (LinColor output of a plain solid paint on canvases the fused small-canvas kernel takes): a share of every
    // chunk is rendered as coverage (4 B per pixel over PCIe instead of 16) and turned into colour * alpha by host threads
    // while the rest of the chunk arrives as LinColor by DMA — two producers into the caller's buffer instead of one PCIe
    // link.  The share adapts from call to call (ctx->expand_frac); RGPU_E2E_EXPAND=0 switches it off.
    const bool expand_enabled = !(getenv("RGPU_E2E_EXPAND") && atoi(getenv("RGPU_E2E_EXPAND")) == 0);  // read per call: A/B inside one process
    const bool can_expand = expand_enabled && out_format == RGPU_OUT_LINCOLOR && plain_solid(paint) && width <= 64 && height <= 64 && n_paths >= 64;
    static const char* fixed_share = getenv("RGPU_E2E_EXPAND_FRAC");  // diagnosis: a fixed share instead of the adaptive one
    if (fixed_share) ctx->expand_frac = std::min(1.0, std::max(0.0, atof(fixed_share)));
    for (int i = 0; i < ring; i++) {
        if ((rc = ensure_dev(ctx, ctx->ring_slab[i], chunk * px * slab_elem))) return rc;
        if ((out_format == RGPU_OUT_RGBA8 || can_expand) && (rc = ensure_dev(ctx, ctx->ring_rgba[i], chunk * px * 4))) return rc;
        if (can_expand && (rc = ensure_pinned(ctx, ctx->h_alpha[i], ctx->h_alpha_cap[i], chunk * px))) return rc;
    }
    if (can_expand) ensure_pool(ctx);
    float colour[4] = {0.f, 0.f, 0.f, 0.f};
    if (can_expand) std::memcpy(colour, paint->solid, sizeof(colour));
    double wait_pool_ms = 0.0, prep_ms = 0.0, submit_ms = 0.0;
    auto ms_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    // The expansion tasks of a chunk are handed to the pool as soon as its copies are queued; each task first sleeps on the
    // event behind the coverage copy (the pool is first in, first out, so the threads take the chunks in copy order) — the
    // submitting thread never waits for the pool except for a staging slot: the coverage staging of a slot may be overwritten
    // once the tasks of the chunk that used it last (`ring` chunks earlier) are done.
    std::atomic<int> exp_left[rgpu_ctx::kRing];
    for (auto& e : exp_left) e.store(0);
    struct PoolDrain {  // no task may outlive exp_left, whichever way the call ends
        rgpu::HostPool* p;
        ~PoolDrain() { if (p) p->wait(); }
    } pool_drain{can_expand ? ctx->pool.get() : nullptr};
    auto wait_slot_expanded = [&](int slot) {
        if (exp_left[slot].load(std::memory_order_acquire) == 0) return;
        const auto t0 = std::chrono::steady_clock::now();
        while (exp_left[slot].load(std::memory_order_acquire) > 0) std::this_thread::yield();
        wait_pool_ms += ms_since(t0);
    };
    // The control points go up chunk by chunk, next to the chunk's items (both are small against the chunk's output): one
    // copy of all of them up front — 115 MB for 100 000 glyphs, from the caller's pageable array — held the first kernel
    // back by ~10 ms of the 130 ms call, while a chunk's share overlaps the download of the chunk before it.
    std::vector<uint32_t> pt_off;
    segment_point_offsets(all, pt_off);
    ctx->staged_items_valid = false;
    auto first_point = [&](size_t path) -> uint32_t {
        const uint32_t sp = pso[path];
        return pt_off[sp < all->n_subpaths ? all->subpath_offsets[sp] : all->n_segments];
    };
    static const double ident[6] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0};
    // Host preparation of a chunk — its item lists, and its control points moved into pinned staging (an asynchronous DMA
    // instead of the driver's synchronous bounce of pageable memory) — runs on a helper thread one chunk ahead of the
    // submissions: ~0.8 ms per chunk of 2048 glyphs, 40 ms of a 100 000-glyph call when it sat between the submissions.
    // Two staging slots; a slot is released after its chunk's submission has synchronised the stream (its copies are done).
    size_t max_pts = 1, max_items = 1;
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t a = c * chunk, b = std::min(n_paths, a + chunk);
        max_pts = std::max<size_t>(max_pts, first_point(b) - first_point(a));
        const uint32_t sa = pso[a], sb = pso[b];
        max_items = std::max<size_t>(max_items, (size_t)(all->subpath_offsets[sb] - all->subpath_offsets[sa]) + (sb - sa));
    }
    const size_t stage_items_at = (sizeof(double2) * max_pts + 255) & ~(size_t)255;
    for (int i = 0; i < 2; i++)
        if ((rc = ensure_pinned(ctx, ctx->h_chunk[i], ctx->h_chunk_cap[i], stage_items_at + sizeof(uint2) * 2 * max_items))) return rc;
    if ((rc = ensure_dev(ctx, ctx->tmp_pts, std::max<size_t>(sizeof(double2) * max_pts, 16)))) return rc;
    if ((rc = ensure_dev(ctx, ctx->tmp_items, sizeof(uint2) * 2 * max_items))) return rc;
    struct Prep {
        std::vector<uint2> ref, packed;
        std::vector<uint32_t> item_off, n_curves;
        uint32_t pt_a = 0, pt_b = 0;
    } preps[2];
    auto prepare = [&](size_t c) {
        Prep& pr = preps[c & 1];
        const size_t a = c * chunk, b = std::min(n_paths, a + chunk);
        pr.ref.clear();
        pr.packed.clear();
        pr.pt_a = first_point(a);
        pr.pt_b = first_point(b);
        build_range_items(all, pso, pt_off, a, b, pr.pt_a, pr.ref, pr.packed, pr.item_off, pr.n_curves);
        unsigned char* const st = ctx->h_chunk[c & 1];
        if (pr.pt_b > pr.pt_a) std::memcpy(st, all->points + 2 * (size_t)pr.pt_a, sizeof(double2) * (pr.pt_b - pr.pt_a));
        std::memcpy(st + stage_items_at, pr.ref.data(), sizeof(uint2) * pr.ref.size());
        std::memcpy(st + stage_items_at + sizeof(uint2) * pr.ref.size(), pr.packed.data(), sizeof(uint2) * pr.packed.size());
    };
    struct Ahead {  // the helper thread and its handshake; the destructor stops and joins it on every way out
        std::mutex m;
        std::condition_variable cv;
        size_t produced = 0, consumed = 0;
        bool stop = false;
        std::thread th;
        ~Ahead() {
            {
                std::lock_guard<std::mutex> l(m);
                stop = true;
            }
            cv.notify_all();
            if (th.joinable()) th.join();
        }
    } ahead;
    if (n_chunks > 1)
        ahead.th = std::thread([&] {
            for (size_t c = 0; c < n_chunks; c++) {
                {
                    std::unique_lock<std::mutex> l(ahead.m);
                    ahead.cv.wait(l, [&] { return ahead.stop || c < ahead.consumed + 2; });
                    if (ahead.stop) return;
                }
                prepare(c);
                {
                    std::lock_guard<std::mutex> l(ahead.m);
                    ahead.produced = c + 1;
                }
                ahead.cv.notify_all();
            }
        });
    std::vector<rgpu_dpath> dps(chunk);
    std::vector<rgpu_job> jobs(chunk);
    int status = RGPU_OK;
    for (size_t c = 0; c < n_chunks && status == RGPU_OK; c++) {
        const size_t a = c * chunk, b = std::min(n_paths, a + chunk), n = b - a;
        const int slot = (int)(c % ring);
        const auto t_prep = std::chrono::steady_clock::now();
        if (n_chunks > 1) {
            std::unique_lock<std::mutex> l(ahead.m);
            ahead.cv.wait(l, [&] { return ahead.produced > c; });
        } else {
            prepare(c);
        }
        const Prep& pr = preps[c & 1];
        const std::vector<uint32_t>&item_off = pr.item_off, &n_curves = pr.n_curves;
        const uint32_t pt_a = pr.pt_a, pt_b = pr.pt_b;
        const size_t n_items = pr.ref.size();
        const unsigned char* const st = ctx->h_chunk[c & 1];
        double2* const d_pts = static_cast<double2*>(ctx->tmp_pts.p);
        uint2* const d_items = static_cast<uint2*>(ctx->tmp_items.p);
        // (the previous chunk's kernel has completed — submit_sync — so its points and items may be overwritten)
        if (pt_b > pt_a) CK(ctx, cudaMemcpyAsync(d_pts, st, sizeof(double2) * (pt_b - pt_a), cudaMemcpyHostToDevice, ctx->stream));
        if (n_items) CK(ctx, cudaMemcpyAsync(d_items, st + stage_items_at, sizeof(uint2) * 2 * n_items, cudaMemcpyHostToDevice, ctx->stream));
        ctx->last_h2d_bytes += sizeof(double2) * (pt_b - pt_a) + sizeof(uint2) * 2 * n_items;
        char* const slab = static_cast<char*>(ctx->ring_slab[slot].p);
        // paths [a, a + n_dma) arrive as LinColor by DMA, paths [a + n_dma, b) as coverage, expanded on the host
        const size_t n_exp = (can_expand && n >= 32) ? std::min(n - 1, (size_t)((double)n * ctx->expand_frac)) : 0;
        const size_t n_dma = n - n_exp;
        size_t live = 0;
        for (size_t i = 0; i < n; i++) {
            rgpu_dpath& d = dps[i];
            d.pts = d_pts;
            d.items = d_items + item_off[i];
            d.items_packed = d_items + n_items + item_off[i];
            d.n_items = item_off[i + 1] - item_off[i];
            d.n_curves = n_curves[i];
            d.n_points = pt_b - pt_a;
            rgpu_job& j = jobs[i];
            std::memset(&j, 0, sizeof(j));
            j.path = &d;
            std::memcpy(j.tr, trs ? trs + 6 * (a + i) : ident, sizeof(j.tr));
            j.fill_rule = fill_rule;
            const bool as_cov = coverage || i >= n_dma;
            j.mode = as_cov ? RGPU_JOB_COVERAGE : RGPU_JOB_RENDER;
            j.paint = as_cov ? nullptr : paint;
            j.canvas = (i >= n_dma && !coverage) ? ctx->ring_rgba[slot].p : slab;
            j.origin = (i >= n_dma && !coverage) ? (i - n_dma) * px : i * px;
            j.row_stride = width;
            j.width = width;
            j.height = height;
            live += d.n_items != 0;
        }
        // the slab is free once the download of the chunk that used it last has finished
        if (c >= (size_t)ring) CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ring_copied[slot], 0));
        if (live != n) {  // paths without segments draw nothing: their images are the fresh (zero) canvas
            CK(ctx, cudaMemsetAsync(slab, 0, n_dma * px * slab_elem, ctx->stream));
            if (n_exp) CK(ctx, cudaMemsetAsync(ctx->ring_rgba[slot].p, 0, n_exp * px * 4, ctx->stream));
        }
        prep_ms += ms_since(t_prep);
        const auto t_submit = std::chrono::steady_clock::now();
        status = submit_sync(ctx, jobs.data(), n, RGPU_BATCH_INDEPENDENT, 1);
        submit_ms += ms_since(t_submit);
        if (n_chunks > 1) {  // the stream is synchronised: the staging slot of this chunk is free for chunk c + 2
            {
                std::lock_guard<std::mutex> l(ahead.m);
                ahead.consumed = c + 1;
            }
            ahead.cv.notify_all();
        }
        if (status != RGPU_OK) break;
        const void* src = slab;
        if (out_format == RGPU_OUT_RGBA8) {
            launch_to_rgba8(reinterpret_cast<const float4*>(slab), static_cast<uchar4*>(ctx->ring_rgba[slot].p), n * px, ctx->stream);
            ctx->n_launches++;
            src = ctx->ring_rgba[slot].p;
        }
        CK(ctx, cudaEventRecord(ctx->ring_done[slot], ctx->stream));
        CK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ring_done[slot], 0));
        // The coverage share first, in pieces (RGPU_E2E_PIECE_MB, default 2 MB; 0: the share in one copy): each piece has its own
        // event and its own expansion task, so a piece is multiplied out right after it has landed — while it is still in the
        // last-level cache the DMA wrote it into, instead of coming back from DRAM after the whole share (30 MB) has arrived.
        size_t n_pieces = 0, piece_px = 0;
        if (n_exp) {
            static const size_t piece_mb = getenv("RGPU_E2E_PIECE_MB") ? (size_t)std::max(0, atoi(getenv("RGPU_E2E_PIECE_MB"))) : 2;
            const size_t total = n_exp * px;
            piece_px = piece_mb ? std::max<size_t>(4096, (piece_mb << 20) / 4) : total;
            n_pieces = (total + piece_px - 1) / piece_px;
            std::vector<cudaEvent_t>& evs = ctx->ring_alpha[slot];
            while (evs.size() < n_pieces) {
                cudaEvent_t e;
                CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
                evs.push_back(e);
            }
            wait_slot_expanded(slot);
            const float* d_alpha = static_cast<const float*>(ctx->ring_rgba[slot].p);
            for (size_t q = 0; q < n_pieces; q++) {
                const size_t lo = q * piece_px, hi = std::min(total, lo + piece_px);
                CK(ctx, cudaMemcpyAsync(ctx->h_alpha[slot] + lo, d_alpha + lo, (hi - lo) * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
                CK(ctx, cudaEventRecord(evs[q], ctx->copy_stream));
            }
        }
        CK(ctx, cudaMemcpyAsync(static_cast<char*>(out_host) + a * px * out_elem, src, n_dma * px * out_elem, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CK(ctx, cudaEventRecord(ctx->ring_copied[slot], ctx->copy_stream));
        ctx->last_d2h_bytes += n_dma * px * out_elem + n_exp * px * 4;
        if (n_exp) {
            const float* alpha = ctx->h_alpha[slot];
            float* dst = static_cast<float*>(out_host) + (a + n_dma) * px * 4;
            const size_t total = n_exp * px;
            // a piece is one task, or — when the share is one or two pieces — enough tasks for the whole pool
            const size_t per_piece = n_pieces >= ctx->pool->size() ? 1 : (2 * ctx->pool->size() + n_pieces - 1) / n_pieces;
            std::atomic<int>* const left = &exp_left[slot];
            left->store((int)(n_pieces * per_piece), std::memory_order_release);
            for (size_t q = 0; q < n_pieces; q++) {
                const size_t plo = q * piece_px, phi = std::min(total, plo + piece_px);
                const cudaEvent_t landed = ctx->ring_alpha[slot][q];
                for (size_t t = 0; t < per_piece; t++) {
                    const size_t lo = plo + (phi - plo) * t / per_piece, hi = plo + (phi - plo) * (t + 1) / per_piece;
                    ctx->pool->submit([=] {
                        cudaEventSynchronize(landed);
                        if (hi > lo) expand_alpha(alpha + lo, colour, dst + 4 * lo, hi - lo);
                        left->fetch_sub(1, std::memory_order_release);
                    });
                }
            }
        }
    }
    if (ctx->pool && can_expand) {
        const auto t0 = std::chrono::steady_clock::now();
        ctx->pool->wait();
        wait_pool_ms += ms_since(t0);
    }
    static const bool trace = getenv("RGPU_E2E_TRACE") != nullptr;
    const auto t_tail = std::chrono::steady_clock::now();
    cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (can_expand && !fixed_share && n_chunks >= 4 && status == RGPU_OK) {
        // The share is picked from the calls' own times: who waits for whom — copy engine, expansion threads, host memory —
        // shows up only in the total.  Eleven candidate shares (0.0, 0.1 .. 1.0: several GPUs
        // behind one host's memory are best served by plain DMA, one GPU by expanding nearly everything); a call's time per pixel goes into its share's
        // running mean, the next call takes the best share seen so far, or a neighbour of it that has not been tried yet.
        // (The first call of a context pays for the pinned allocations and is not recorded; another workload starts afresh.)
        const double key = (double)n_paths * (double)px, per_px = ms_since(t_call) / key;
        if (!(key > 0.75 * ctx->share_key && key < 1.33 * ctx->share_key)) {
            std::fill(ctx->share_ms, ctx->share_ms + rgpu_ctx::kShares, 0.0);
            ctx->share_key = key;
        }
        if (ctx->share_cold) ctx->share_cold = false;
        else {
            double& m = ctx->share_ms[ctx->share_cur];
            m = m > 0.0 ? 0.6 * m + 0.4 * per_px : per_px;
        }
        int best = ctx->share_cur;
        for (int i = 0; i < rgpu_ctx::kShares; i++)
            if (ctx->share_ms[i] > 0.0 && (ctx->share_ms[best] <= 0.0 || ctx->share_ms[i] < ctx->share_ms[best])) best = i;
        if (ctx->share_ms[best] > 0.0) {
            if (best + 1 < rgpu_ctx::kShares && ctx->share_ms[best + 1] <= 0.0) best++;
            else if (best > 0 && ctx->share_ms[best - 1] <= 0.0) best--;
        }
        ctx->share_cur = best;
        ctx->expand_frac = 0.1 * best;
    }
    if (trace)
        fprintf(stderr, "rgpu_fill_batch_host: %zu chunks, waited %.2f ms for prepared chunks, %.2f ms in submit + status (kernels, and the slab's "
                "previous download), %.2f ms for the expansion threads, %.2f ms for the last copies; next share %.2f\n", n_chunks, prep_ms, submit_ms,
                wait_pool_ms, ms_since(t_tail), can_expand ? ctx->expand_frac : 0.0);
    if (status != RGPU_OK) return status;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        ctx->err = std::string("rgpu_fill_batch_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);
        return RGPU_ERR_CUDA;
    }
    return RGPU_OK;
}

int rgpu_mask_banded_host(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, void* img, size_t elem_size, size_t width,
                          size_t height, uint32_t n_bands, uint32_t band_first, uint32_t band_count) {
    if (!ctx || !tr) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (elem_size != 4 && elem_size != 8) return fail(ctx, RGPU_ERR_INVALID, "elem_size must be 4 (f32) or 8 (f64)");
    if (fill_rule != RGPU_NONZERO && fill_rule != RGPU_EVENODD) return fail(ctx, RGPU_ERR_INVALID, "bad fill rule");
    if (width > 0x7ffffff0u || height > 0x7ffffff0u) return fail(ctx, RGPU_ERR_INVALID, "image too large");
    ctx->last_h2d_bytes = ctx->last_d2h_bytes = 0;
    if (width == 0 || height == 0) return RGPU_OK;
    if (!img) return RGPU_ERR_INVALID;
    int rc = validate_path(ctx, path);
    if (rc) return rc;
    if (n_bands == 0) n_bands = 1;
    n_bands = (uint32_t)std::min<size_t>(n_bands, (height + 7) / 8);
    auto cut = [&](uint32_t k) -> size_t {  // multiples of the raster tile height, like sharding.band_rows
        if (k >= n_bands) return height;
        const size_t y = height * k / n_bands;
        return std::min(height, (y + 7) / 8 * 8);
    };
    struct Band { size_t y0, rows, off; };
    std::vector<Band> bands;
    size_t rows_total = 0;
    for (uint32_t b = band_first; b < n_bands && b - band_first < band_count; b++) {
        const size_t y0 = cut(b), y1 = std::max(cut(b), cut(b + 1));
        if (y1 > y0) {
            // consecutive bands are one job (flattened once)
            if (!bands.empty() && bands.back().y0 + bands.back().rows == y0) bands.back().rows += y1 - y0;
            else bands.push_back({y0, y1 - y0, rows_total});
            rows_total += y1 - y0;
        }
    }
    if (bands.empty()) return RGPU_OK;
    if ((rc = ensure_dev(ctx, ctx->img_f32, sizeof(float) * width * rows_total))) return rc;
    float* const d_img = static_cast<float*>(ctx->img_f32.p);
    const bool empty = path->n_segments == 0 || path->n_subpaths == 0;
    if (empty) {
        CK(ctx, cudaMemsetAsync(d_img, 0, sizeof(float) * width * rows_total, ctx->stream));
    } else {
        rgpu_dpath dp;
        if ((rc = stage_path(ctx, path, &dp))) return rc;
        std::vector<rgpu_job> jobs(bands.size());
        std::vector<double> origins(bands.size());
        for (size_t i = 0; i < bands.size(); i++) {
            rgpu_job& j = jobs[i];
            std::memset(&j, 0, sizeof(j));
            j.path = &dp;
            std::memcpy(j.tr, tr, sizeof(j.tr));
            origins[i] = (double)bands[i].y0;  // the band's rows: flattened on the canvas, shifted as lines (JobDev::y_org), cropped by the reference's own y clipping
            j.fill_rule = fill_rule;
            j.mode = RGPU_JOB_MASK;
            j.canvas = d_img;
            j.origin = bands[i].off * width;
            j.row_stride = width;
            j.width = (uint32_t)width;
            j.height = (uint32_t)bands[i].rows;
        }
        ctx->job_row_origin = origins.data();
        rc = submit_sync(ctx, jobs.data(), jobs.size(), RGPU_BATCH_INDEPENDENT, 1);
        ctx->job_row_origin = nullptr;
        if (rc) return rc;
    }
    const uint64_t h2d = ctx->last_h2d_bytes;
    uint64_t d2h = 0;
    {   // large masks come down run-coded (constant segments as class bytes, rebuilt by host threads): see download_runcoded
        std::vector<size_t> img_row(rows_total);
        for (const Band& b : bands)
            for (size_t i = 0; i < b.rows; i++) img_row[b.off + i] = b.y0 + i;
        rc = elem_size == 4 ? download_runcoded<float>(ctx, d_img, width, rows_total, img_row, static_cast<float*>(img), width)
                            : download_runcoded<double>(ctx, d_img, width, rows_total, img_row, static_cast<double*>(img), width);
        if (rc < 0) return rc;
        if (rc == 0) {
            ctx->last_h2d_bytes = h2d;
            return RGPU_OK;
        }
    }
    for (const Band& b : bands) {
        if (elem_size == 4) {
            CK(ctx, cudaMemcpyAsync(static_cast<float*>(img) + b.y0 * width, d_img + b.off * width, sizeof(float) * width * b.rows, cudaMemcpyDeviceToHost, ctx->stream));
            d2h += sizeof(float) * width * b.rows;
        } else {
            rgpu_shape sh{0, width, b.rows, width, 1};
            if ((rc = download_widen(ctx, d_img + b.off * width, width, b.rows, static_cast<double*>(img) + b.y0 * width, sh))) return rc;
            d2h += ctx->last_d2h_bytes;
        }
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->last_h2d_bytes = h2d;
    ctx->last_d2h_bytes = d2h;
    return RGPU_OK;
}

}  // extern "C"

// ---- several GPUs -----------------------------------------------------------------------------------------------
struct rgpu_multi {
    std::vector<rgpu_ctx*> ctxs;
    std::vector<int> devices;
    std::string err;
};

namespace {

thread_local std::string g_multi_err;

// run f(d) for every device on its own host thread; returns the first non-zero status (and its text in m->err)
template <class F>
int multi_run(rgpu_multi* m, F f) {
    const size_t n = m->ctxs.size();
    std::vector<int> rcs(n, RGPU_OK);
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t d = 1; d < n; d++) th.emplace_back([&, d] { rcs[d] = f(d); });
    rcs[0] = f(0);
    for (auto& t : th) t.join();
    for (size_t d = 0; d < n; d++)
        if (rcs[d] != RGPU_OK) {
            m->err = "device " + std::to_string(m->devices[d]) + ": " + m->ctxs[d]->err;
            return rcs[d];
        }
    return RGPU_OK;
}

}  // namespace

extern "C" {

int rgpu_multi_create(const int* devices, int n_devices, double flatness, rgpu_multi** out) {
    if (!out) return RGPU_ERR_INVALID;
    *out = nullptr;
    if (n_devices <= 0) {
        g_multi_err = "n_devices must be positive";
        return RGPU_ERR_INVALID;
    }
    auto* m = new rgpu_multi();
    for (int i = 0; i < n_devices; i++) {
        const int dev = devices ? devices[i] : i;
        rgpu_ctx* c = nullptr;
        int rc = rgpu_create(dev, flatness, &c);
        if (rc != RGPU_OK) {
            g_multi_err = "device " + std::to_string(dev) + ": " + rgpu_last_error(nullptr);
            for (rgpu_ctx* p : m->ctxs) rgpu_destroy(p);
            delete m;
            return rc;
        }
        m->ctxs.push_back(c);
        m->devices.push_back(dev);
    }
    for (rgpu_ctx* c : m->ctxs) c->pool_share = (unsigned)n_devices;  // the contexts' host threads share the cores
    *out = m;
    return RGPU_OK;
}

void rgpu_multi_destroy(rgpu_multi* m) {
    if (!m) return;
    for (rgpu_ctx* c : m->ctxs) rgpu_destroy(c);
    delete m;
}

int rgpu_multi_device_count(const rgpu_multi* m) { return m ? (int)m->ctxs.size() : 0; }

const char* rgpu_multi_last_error(const rgpu_multi* m) { return m ? m->err.c_str() : g_multi_err.c_str(); }

int rgpu_multi_fill_batch_host(rgpu_multi* m, const rgpu_path* all, const uint32_t* pso, size_t n_paths, const double* trs, int fill_rule,
                               const rgpu_paint* paint, uint32_t width, uint32_t height, int out_format, void* out_host) {
    if (!m || m->ctxs.empty()) return RGPU_ERR_INVALID;
    int rc = validate_batch(m->ctxs[0], all, pso, n_paths);
    if (rc) {
        m->err = m->ctxs[0]->err;
        return rc;
    }
    if (out_format < RGPU_OUT_LINCOLOR || out_format > RGPU_OUT_COVERAGE) {
        m->err = "unknown output format";
        return RGPU_ERR_INVALID;
    }
    if (n_paths == 0) return RGPU_OK;
    const size_t nd = m->ctxs.size();
    // contiguous ranges of paths holding equal numbers of segments (balance by segments, not by count)
    std::vector<size_t> cuts(nd + 1, n_paths);
    cuts[0] = 0;
    if (all->n_subpaths == 0) {
        for (size_t d = 1; d < nd; d++) cuts[d] = n_paths * d / nd;
    } else {
        const uint64_t total = all->subpath_offsets[all->n_subpaths];
        size_t i = 0;
        for (size_t d = 1; d < nd; d++) {
            const uint64_t want = total * d / nd;
            while (i < n_paths && (uint64_t)all->subpath_offsets[pso[i]] < want) i++;
            cuts[d] = i;
        }
    }
    // first segment / point of every range
    std::vector<uint32_t> seg0(nd + 1), pt0(nd + 1);
    {
        uint32_t acc = 0, seg = 0;
        for (size_t d = 0; d <= nd; d++) {
            const uint32_t s = all->n_subpaths == 0 ? 0u : all->subpath_offsets[cuts[d] < n_paths ? pso[cuts[d]] : all->n_subpaths];
            for (; seg < s; seg++) acc += all->kinds[seg];
            seg0[d] = s;
            pt0[d] = acc;
        }
    }
    const size_t img_bytes = (size_t)width * height * out_elem_bytes(out_format);
    return multi_run(m, [&](size_t d) -> int {
        const size_t a = cuts[d], b = cuts[d + 1];
        if (b <= a) return RGPU_OK;
        // rebased view of paths [a, b)
        const uint32_t sub0 = pso[a], sub1 = pso[b];
        std::vector<uint32_t> so(sub1 - sub0 + 1), po(b - a + 1);
        for (uint32_t s = sub0; s <= sub1; s++) so[s - sub0] = all->subpath_offsets[s] - seg0[d];
        for (size_t i = a; i <= b; i++) po[i - a] = pso[i] - sub0;
        rgpu_path v;
        v.points = all->points + 2 * (size_t)pt0[d];
        v.kinds = all->kinds + seg0[d];
        v.subpath_offsets = so.data();
        v.closed = all->closed + sub0;
        v.n_points = pt0[d + 1] - pt0[d];
        v.n_segments = seg0[d + 1] - seg0[d];
        v.n_subpaths = sub1 - sub0;
        return rgpu_fill_batch_host(m->ctxs[d], &v, po.data(), b - a, trs ? trs + 6 * a : nullptr, fill_rule, paint, width, height, out_format,
                                    static_cast<char*>(out_host) + a * img_bytes);
    });
}

int rgpu_multi_mask_banded_host(rgpu_multi* m, const rgpu_path* path, const double tr[6], int fill_rule, void* img, size_t elem_size, size_t width,
                                size_t height, uint32_t n_bands) {
    if (!m || m->ctxs.empty() || !tr) return RGPU_ERR_INVALID;
    const uint32_t nd = (uint32_t)m->ctxs.size();
    if (n_bands == 0) n_bands = nd;
    n_bands = (uint32_t)std::min<size_t>(n_bands, std::max<size_t>((height + 7) / 8, 1));  // the clamp rgpu_mask_banded_host applies
    return multi_run(m, [&](size_t d) -> int {
        const uint32_t b0 = (uint32_t)((uint64_t)n_bands * d / nd), b1 = (uint32_t)((uint64_t)n_bands * (d + 1) / nd);
        return rgpu_mask_banded_host(m->ctxs[d], path, tr, fill_rule, img, elem_size, width, height, n_bands, b0, b1 - b0);
    });
}

}  // extern "C"
