// C ABI of rasterize_b200 (include/rasterize_b200.h): context, scratch management, job submission and the
// host-buffer (trait-level) entry points.  No CPU fallback lives here: every entry point either launches the
// CUDA kernels or fails.
#include "../../include/rasterize_b200.h"
#include "rgpu_internal.cuh"
#include "host_pool.hpp"

#include <algorithm>
#include <atomic>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace rgpu;

struct rgpu_dpath {
    double2* pts = nullptr;
    uint2* items = nullptr;         // reference order (ordered flatten, two-pass binning)
    uint2* items_packed = nullptr;  // curves first, then lines / closing items (single-pass binning)
    uint32_t n_points = 0, n_items = 0, n_curves = 0;
    uint32_t n_segments = 0, n_subpaths = 0;  // n_items = n_segments + n_subpaths
};

namespace {

thread_local std::string g_create_err;

constexpr uint64_t kFixedBinBudget = 2ull << 30;  // bytes of fixed-capacity tile bins before the two-pass scheme takes over

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct HostXf {  // 2x3 affine, reference src/geometry.rs:317-384, 519-539 (host-side paint set-up only)
    double m[6];
    static HostXf from(const double* t) { HostXf r; std::memcpy(r.m, t, sizeof(r.m)); return r; }
    HostXf mul(const HostXf& o) const {
        HostXf r;
        r.m[0] = m[0] * o.m[0] + m[1] * o.m[3];
        r.m[1] = m[0] * o.m[1] + m[1] * o.m[4];
        r.m[2] = m[0] * o.m[2] + m[1] * o.m[5] + m[2];
        r.m[3] = m[3] * o.m[0] + m[4] * o.m[3];
        r.m[4] = m[3] * o.m[1] + m[4] * o.m[4];
        r.m[5] = m[3] * o.m[2] + m[4] * o.m[5] + m[5];
        return r;
    }
    bool invert(HostXf& out) const {
        double det = m[0] * m[4] - m[3] * m[1];
        if (std::fabs(det) <= 2.220446049250313e-16) return false;
        double o00 = m[4] / det, o01 = -m[1] / det, o10 = -m[3] / det, o11 = m[0] / det;
        out.m[0] = o00; out.m[1] = o01; out.m[2] = -o00 * m[2] - o01 * m[5];
        out.m[3] = o10; out.m[4] = o11; out.m[5] = -o10 * m[2] - o11 * m[5];
        return true;
    }
};

}  // namespace

struct rgpu_ctx {
    int device = 0;
    double flatness = 0.05;
    cudaStream_t stream = nullptr;
    std::string err;
    // device scratch (grow-only)
    DevBuf jobs, paints, slot_counts, slot_offs, lines, line_job, zero_block, tile_offs, refs, scan_temp, status, tile_state;
    uint32_t epoch = 0;
    int fix_shift = kFixShift;  // fraction bits of the winding cells of the batch being submitted (see rgpu_internal.cuh)
    const double* job_row_origin = nullptr;  // per job of the batch being submitted: JobDev::y_org (set by rgpu_mask_banded_host around its submission)
    DevBuf img_f32, img_f64, img_lin;  // staging canvases of the host-buffer entry points
    DevBuf px_counts, px_out;          // rgpu_mask_iter: per-block pixel counts | offsets, compacted records
    DevBuf rc_buf, rc_lits;            // run-coded download: [row counts | row offsets | class bytes], literals
    unsigned char* h_rc = nullptr;     // pinned: [row offsets | class bytes]
    size_t h_rc_cap = 0;
    float* h_lits = nullptr;           // pinned: literals
    size_t h_lits_cap = 0;
    unsigned pool_share = 1;           // contexts sharing the host's cores (rgpu_multi_create)
    int rc_skip = 0;                   // calls of size rc_skip_w x rc_skip_h that skip the run-coded attempt (declined last time)
    size_t rc_skip_w = 0, rc_skip_h = 0;
    DevBuf stroke_buf;                 // scratch of rgpu_path_stroke (unit table, counts, offsets, first / last pieces)
    DevBuf tmp_pts, tmp_items;         // device copy of the path of the current host-buffer call (grow-only, no per-call cudaMalloc)
    uint2* h_items = nullptr;          // pinned staging of the item list
    size_t h_items_cap = 0;
    double2* h_pts = nullptr;          // pinned staging of the control points of the current host-buffer call
    size_t h_pts_cap = 0;
    // item lists staged by the last single-path host-buffer call, valid while tmp_items holds them (see stage_path)
    std::vector<unsigned char> staged_meta, staged_meta_tmp;
    rgpu_dpath staged_dp;
    bool staged_items_valid = false;
    size_t lines_cap = 0, refs_cap = 0;
    // fixed-capacity tile bins (single flatten walk): bin_cap is the largest per-tile capacity any batch needed so
    // far; a batch whose tiles x capacity exceeds kFixedBinBudget uses the exact count -> scan -> emit scheme
    uint32_t bin_cap = 0;
    bool two_pass = getenv("RGPU_TWO_PASS") != nullptr;  // A/B switch: always use the exact two-pass scheme
    bool last_fixed = false;
    DevBuf fixed_block;            // [status A | status B | tickets | tile counters], self-cleaning (see submit)
    uint32_t fx_tickets_cap = 0, fx_tiles_cap = 0, fx_parity = 0;
    cudaEvent_t h_tables_ev = nullptr;  // completion of the last upload out of h_jobs / h_paints
    // what the device job / paint tables hold: a batch that is submitted again (the steady state of a render loop) finds
    // its tables already in HBM and uploads nothing
    struct TableShadow {
        std::vector<unsigned char> bytes;
        const void* dev = nullptr;
    } jobs_shadow, paints_shadow, lists_shadow;
    DevBuf scene_lists;                  // per-band job lists of a scene batch (SceneArgs::band_offs / band_jobs)
    std::vector<uint32_t> h_scene_lists;
    uint32_t last_tiles = 0;
    // pinned host
    Status* h_status = nullptr;
    void* h_stage = nullptr;
    size_t h_stage_cap = 0;
    JobDev* h_jobs = nullptr;
    size_t h_jobs_cap = 0;
    PaintDev* h_paints = nullptr;
    size_t h_paints_cap = 0;
    // stats
    uint64_t n_launches = 0, last_lines = 0, last_refs = 0;
    uint64_t last_h2d_bytes = 0, last_d2h_bytes = 0;  // PCIe traffic of the last host-buffer call
    // last submission (for status / retry)
    uint64_t need_lines = 0, need_refs = 0;
    uint32_t last_total_slots = 0;
    Status* d_status_cur = nullptr;  // device status block of the last submission, not yet fetched
    // host-side result pipeline (chunked D2H overlapped with threaded widening into the caller's image)
    std::unique_ptr<rgpu::HostPool> pool;
    std::vector<cudaEvent_t> chunk_ev;
    double widen_dev_frac = 0.2;  // share of the rows of an f64 result widened on the device (adapts, see download_widen)
    // rgpu_fill_batch_host: second stream for the downloads, a ring of output slabs and their events
    static constexpr int kRing = 3;
    cudaStream_t copy_stream = nullptr;
    DevBuf ring_slab[kRing], ring_rgba[kRing];
    unsigned char* h_chunk[2] = {nullptr, nullptr};       // pinned staging of a chunk's control points + item lists (rgpu_fill_batch_host)
    size_t h_chunk_cap[2] = {0, 0};
    float* h_alpha[kRing] = {nullptr, nullptr, nullptr};  // pinned staging of the coverage share of a chunk (split download)
    size_t h_alpha_cap[kRing] = {0, 0, 0};
    static constexpr int kShares = 11;
    double expand_frac = 0.8;   // share of a chunk's images that crosses PCIe as coverage and is expanded by host threads (adapts)
    double share_ms[kShares] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // time per pixel of the calls at share 0.1 i (running mean; 0: not tried)
    double share_key = 0.0;     // pixels of the workload the table belongs to
    int share_cur = 8;
    bool share_cold = true;
    cudaEvent_t ring_done[kRing] = {nullptr, nullptr, nullptr}, ring_copied[kRing] = {nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> ring_alpha[kRing];  // piece q of a chunk's coverage share has landed in h_alpha (blocking sync: pool threads sleep on it)
    // optional stage timing
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
};

namespace {

#define CK(ctx, call)                                                                                       \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
            return RGPU_ERR_CUDA;                                                                           \
        }                                                                                                   \
    } while (0)

int fail(rgpu_ctx* ctx, int code, const char* msg) {
    ctx->err = msg;
    return code;
}

int ensure_dev(rgpu_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return RGPU_OK;
    size_t want = std::max(bytes, b.cap + b.cap / 2);
    want = (want + 255) & ~(size_t)255;
    if (b.p) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    CK(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return RGPU_OK;
}

template <class T>
int ensure_pinned(rgpu_ctx* ctx, T*& p, size_t& cap, size_t count) {
    if (count <= cap) return RGPU_OK;
    size_t want = std::max(count, cap + cap / 2);
    if (p) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaFreeHost(p));
        p = nullptr;
        cap = 0;
    }
    CK(ctx, cudaMallocHost(reinterpret_cast<void**>(&p), want * sizeof(T)));
    cap = want;
    return RGPU_OK;
}

// Copy a host table to its device buffer unless the buffer already holds exactly these bytes.
int upload_table(rgpu_ctx* ctx, rgpu_ctx::TableShadow& sh, void* dev, const void* host, size_t bytes, bool& uploaded) {
    if (sh.dev == dev && sh.bytes.size() == bytes && std::memcmp(sh.bytes.data(), host, bytes) == 0) return RGPU_OK;
    CK(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    sh.bytes.assign(static_cast<const unsigned char*>(host), static_cast<const unsigned char*>(host) + bytes);
    sh.dev = dev;
    uploaded = true;
    return RGPU_OK;
}

// The context's host thread pool (widening, expansion and rebuilding of downloaded results).  RGPU_HOST_THREADS, else the host's
// cores divided by the contexts that share them (rgpu_multi_create sets pool_share), capped at 32.
void ensure_pool(rgpu_ctx* ctx) {
    if (ctx->pool) return;
    unsigned n = std::thread::hardware_concurrency();
    n = n ? n : 4u;
    if (const char* e = getenv("RGPU_HOST_THREADS")) n = (unsigned)std::max(1, atoi(e));
    else if (ctx->pool_share > 1) n = std::max(2u, n / ctx->pool_share);
    ctx->pool.reset(new rgpu::HostPool(std::max(1u, std::min(n, 32u))));
}

int ensure_stage(rgpu_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->h_stage_cap) return RGPU_OK;
    if (ctx->h_stage) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaFreeHost(ctx->h_stage));
        ctx->h_stage = nullptr;
        ctx->h_stage_cap = 0;
    }
    CK(ctx, cudaMallocHost(&ctx->h_stage, bytes));
    ctx->h_stage_cap = bytes;
    return RGPU_OK;
}

int validate_path(rgpu_ctx* ctx, const rgpu_path* p) {
    if (!p) return fail(ctx, RGPU_ERR_INVALID, "path is NULL");
    if (p->n_segments && (!p->points || !p->kinds)) return fail(ctx, RGPU_ERR_INVALID, "path arrays are NULL");
    if (p->n_subpaths && (!p->subpath_offsets || !p->closed)) return fail(ctx, RGPU_ERR_INVALID, "subpath arrays are NULL");
    size_t np = 0;
    for (uint32_t i = 0; i < p->n_segments; i++) {
        uint8_t k = p->kinds[i];
        if (k < 2 || k > 4) return fail(ctx, RGPU_ERR_INVALID, "segment kind must be 2 (line), 3 (quad) or 4 (cubic)");
        np += k;
    }
    if (np != p->n_points) return fail(ctx, RGPU_ERR_INVALID, "n_points does not match the sum of kinds");
    if (p->n_points > kItemIndexMask) return fail(ctx, RGPU_ERR_INVALID, "path too large");
    uint32_t prev = 0;
    for (uint32_t s = 0; s < p->n_subpaths; s++) {
        uint32_t a = p->subpath_offsets[s], b = p->subpath_offsets[s + 1];
        if (a != prev || b <= a || b > p->n_segments) return fail(ctx, RGPU_ERR_INVALID, "bad subpath offsets");
        prev = b;
    }
    return RGPU_OK;
}

// Build the device item list: the segments of each subpath followed by its closing item
// (reference order of PathFlattenIter, src/path.rs:761-795).
void pack_items(const std::vector<uint2>& items, std::vector<uint2>& packed) {
    packed.clear();
    packed.reserve(items.size());
    for (const uint2& it : items)
        if (!(it.y & kItemClosing) && it.y != 2u) packed.push_back(it);
    for (const uint2& it : items)
        if ((it.y & kItemClosing) || it.y == 2u) packed.push_back(it);
}

void build_items(const rgpu_path* p, std::vector<uint2>& items, uint32_t& n_curves) {
    std::vector<uint32_t> pt_off(p->n_segments + 1);
    uint32_t acc = 0;
    n_curves = 0;
    for (uint32_t i = 0; i < p->n_segments; i++) {
        pt_off[i] = acc;
        acc += p->kinds[i];
        if (p->kinds[i] != 2) n_curves++;
    }
    pt_off[p->n_segments] = acc;
    items.clear();
    items.reserve(p->n_segments + p->n_subpaths);
    for (uint32_t s = 0; s < p->n_subpaths; s++) {
        uint32_t a = p->subpath_offsets[s], b = p->subpath_offsets[s + 1];
        for (uint32_t i = a; i < b; i++) items.push_back(make_uint2(pt_off[i], p->kinds[i]));
        uint32_t end_pt = pt_off[b] - 1;  // subpath.end()
        uint32_t start_pt = pt_off[a];    // subpath.start()
        items.push_back(make_uint2(end_pt, kItemClosing | (p->closed[s] ? kItemExplicitClosed : 0u) | start_pt));
    }
}

int upload_path(rgpu_ctx* ctx, const rgpu_path* path, rgpu_dpath* dp) {
    std::vector<uint2> items, packed;
    build_items(path, items, dp->n_curves);
    pack_items(items, packed);
    dp->n_points = path->n_points;
    dp->n_items = (uint32_t)items.size();
    dp->n_segments = path->n_segments;
    dp->n_subpaths = path->n_subpaths;
    items.insert(items.end(), packed.begin(), packed.end());  // one allocation: [reference order | curves first]
    if (dp->n_points) {
        CK(ctx, cudaMalloc(reinterpret_cast<void**>(&dp->pts), sizeof(double2) * dp->n_points));
        CK(ctx, cudaMemcpyAsync(dp->pts, path->points, sizeof(double2) * dp->n_points, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (dp->n_items) {
        CK(ctx, cudaMalloc(reinterpret_cast<void**>(&dp->items), sizeof(uint2) * items.size()));
        CK(ctx, cudaMemcpyAsync(dp->items, items.data(), sizeof(uint2) * items.size(), cudaMemcpyHostToDevice, ctx->stream));
        dp->items_packed = dp->items + dp->n_items;
    }
    // `items` is pageable: the async copy is staged by the driver before returning, but be explicit
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RGPU_OK;
}

// Device copy of a host path in the context's grow-only scratch (host-buffer entry points): no cudaMalloc / cudaFree
// and no extra synchronisation per call; the copies are ordered before the kernels on the context's stream.
// The control points — the call's input — go through pinned staging every time (a true asynchronous DMA instead of the
// driver's synchronous bounce of pageable memory).  The item lists are DERIVED from the path's structure (kinds, subpath
// offsets, closed flags): when that structure is the one staged by the previous call they are already on the device.
int stage_path(rgpu_ctx* ctx, const rgpu_path* path, rgpu_dpath* dp) {
    int rc;
    const size_t pts_bytes = sizeof(double2) * path->n_points;
    if ((rc = ensure_dev(ctx, ctx->tmp_pts, std::max<size_t>(pts_bytes, 16)))) return rc;
    if ((rc = ensure_pinned(ctx, ctx->h_pts, ctx->h_pts_cap, std::max<size_t>(path->n_points, 1)))) return rc;
    // a previous call's copies out of the pinned buffers have completed: every host-buffer entry point ends with a stream sync
    dp->pts = static_cast<double2*>(ctx->tmp_pts.p);
    dp->n_points = path->n_points;
    if (pts_bytes) {
        std::memcpy(ctx->h_pts, path->points, pts_bytes);
        CK(ctx, cudaMemcpyAsync(dp->pts, ctx->h_pts, pts_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->last_h2d_bytes = pts_bytes;
    ctx->last_d2h_bytes = 0;
    // structure of the path: [kinds | subpath offsets | closed flags]
    const uint32_t counts[2] = {path->n_segments, path->n_subpaths};
    const size_t off_bytes = sizeof(uint32_t) * ((size_t)path->n_subpaths + 1);
    std::vector<unsigned char>& meta = ctx->staged_meta_tmp;
    meta.resize(sizeof(counts) + off_bytes + path->n_segments + path->n_subpaths);
    unsigned char* m = meta.data();
    std::memcpy(m, counts, sizeof(counts));
    std::memcpy(m + sizeof(counts), path->subpath_offsets, off_bytes);
    std::memcpy(m + sizeof(counts) + off_bytes, path->kinds, path->n_segments);
    std::memcpy(m + sizeof(counts) + off_bytes + path->n_segments, path->closed, path->n_subpaths);
    if (ctx->staged_items_valid && meta == ctx->staged_meta && ctx->staged_dp.items == static_cast<uint2*>(ctx->tmp_items.p)) {
        dp->items = ctx->staged_dp.items;
        dp->items_packed = ctx->staged_dp.items_packed;
        dp->n_items = ctx->staged_dp.n_items;
        dp->n_curves = ctx->staged_dp.n_curves;
        return RGPU_OK;
    }
    std::vector<uint2> items, packed;
    build_items(path, items, dp->n_curves);
    pack_items(items, packed);
    dp->n_items = (uint32_t)items.size();
    const size_t n2 = items.size() + packed.size();  // [reference order | curves first]
    ctx->staged_items_valid = false;
    if ((rc = ensure_dev(ctx, ctx->tmp_items, sizeof(uint2) * std::max<size_t>(n2, 1)))) return rc;
    if ((rc = ensure_pinned(ctx, ctx->h_items, ctx->h_items_cap, std::max<size_t>(n2, 1)))) return rc;
    std::memcpy(ctx->h_items, items.data(), sizeof(uint2) * items.size());
    std::memcpy(ctx->h_items + items.size(), packed.data(), sizeof(uint2) * packed.size());
    dp->items = static_cast<uint2*>(ctx->tmp_items.p);
    dp->items_packed = dp->items + dp->n_items;
    if (n2) CK(ctx, cudaMemcpyAsync(dp->items, ctx->h_items, sizeof(uint2) * n2, cudaMemcpyHostToDevice, ctx->stream));
    ctx->last_h2d_bytes += sizeof(uint2) * n2;
    ctx->staged_meta.swap(meta);
    ctx->staged_dp = *dp;
    ctx->staged_items_valid = true;
    return RGPU_OK;
}

void free_path(rgpu_dpath* dp) {
    if (dp->pts) cudaFree(dp->pts);
    if (dp->items) cudaFree(dp->items);
    dp->pts = nullptr;
    dp->items = nullptr;
}

// Resolve the paint of a FILL job into the device form; returns false when the fill is a silent no-op
// (singular transform, bounding-box units without a bbox: reference src/rasterize.rs:87-92).
bool build_paint(const rgpu_job& job, PaintDev& out) {
    const rgpu_paint* p = job.paint;
    std::memset(&out, 0, sizeof(out));
    out.kind = p->kind;
    out.linear_colors = p->linear_colors;
    out.spread = p->spread;
    std::memcpy(out.solid, p->solid, sizeof(out.solid));
    if (p->kind == RGPU_PAINT_SOLID) return true;
    HostXf tr = HostXf::from(job.tr);
    HostXf units_tr;
    if (p->units == RGPU_UNITS_USER_SPACE) {
        units_tr = tr.mul(HostXf::from(p->tr));
    } else {
        if (!job.path_bbox) return false;
        const double* bb = job.path_bbox;
        // BBox::unit_transform: translate(x, y).pre_scale(width, height), src/geometry.rs:662-664
        const double tm[6] = {1.0, 0.0, bb[0], 0.0, 1.0, bb[1]};
        const double sm[6] = {bb[2] - bb[0], 0.0, 0.0, 0.0, bb[3] - bb[1], 0.0};
        HostXf t = HostXf::from(tm);
        HostXf s = HostXf::from(sm);
        units_tr = tr.mul(t.mul(s)).mul(HostXf::from(p->tr));
    }
    HostXf inv;
    if (!units_tr.invert(inv)) return false;
    std::memcpy(out.pixel_tr, inv.m, sizeof(inv.m));
    out.p0x = p->p0[0]; out.p0y = p->p0[1]; out.p1x = p->p1[0]; out.p1y = p->p1[1];
    out.r0 = p->r0; out.r1 = p->r1;
    if (p->kind == RGPU_PAINT_LINEAR) {
        // dir = (end - start) / |end - start|^2, src/grad.rs:180-190
        double dx = p->p1[0] - p->p0[0], dy = p->p1[1] - p->p0[1];
        double dd = dx * dx + dy * dy;
        out.dirx = dx / dd;
        out.diry = dy / dd;
    }
    out.n_stops = (int)std::min<uint32_t>(p->n_stops, kMaxStops);
    for (int i = 0; i < out.n_stops; i++) {
        out.stop_pos[i] = p->stop_pos[i];
        std::memcpy(out.stop_col[i], p->stop_colors + 4 * i, 16);
    }
    for (int i = 1; i < out.n_stops; i++) out.stop_inv[i] = 1.0 / (out.stop_pos[i] - out.stop_pos[i - 1]);
    if (p->kind == RGPU_PAINT_LINEAR) {
        // t = (pixel_tr(p) - start) . dir is affine in the pixel centre: fold the two transforms once
        const double* m = out.pixel_tr;
        out.lin_a = m[0] * out.dirx + m[3] * out.diry;
        out.lin_b = m[1] * out.dirx + m[4] * out.diry;
        out.lin_c = (m[2] - out.p0x) * out.dirx + (m[5] - out.p0y) * out.diry;
    } else {
        // the pixel-independent terms of GradRadial::offset, src/grad.rs:361-372 (same expressions, evaluated once)
        out.rad_cdx = out.p0x - out.p1x;
        out.rad_cdy = out.p0y - out.p1y;
        out.rad_rd = out.r0 - out.r1;
        out.rad_a = (out.rad_cdx * out.rad_cdx + out.rad_cdy * out.rad_cdy) - out.rad_rd * out.rad_rd;
        out.rad_inv2a = 1.0 / (2.0 * out.rad_a);
    }
    if (out.n_stops == 0) {  // GradStops::new: empty list -> one opaque black stop, src/grad.rs:92-97
        out.n_stops = 1;
        out.stop_pos[0] = 0.0;
        out.stop_col[0][0] = out.stop_col[0][1] = out.stop_col[0][2] = 0.f;
        out.stop_col[0][3] = 1.f;
    }
    return true;
}

// `scene` (may be NULL): the jobs are the FILL jobs of one dense layer and are composited by the scene kernel (scene.cu) in
// one launch; n_bands / n_chunks of *scene are filled in here.
int submit(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags, int close_flag, bool ordered_lines = false,
           SceneArgs* scene = nullptr);

// Scene batches that cannot use the scene kernel (no fixed bins: two-pass scheme forced or over budget) or have nothing to
// fill: background, the ordered per-fill launches, export.
int scene_fallback(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, int close_flag, const SceneArgs& sc, bool with_jobs) {
    const size_t n = (size_t)sc.width * sc.height;
    if (sc.fresh) {
        launch_fill_color(sc.layer, n, make_float4(sc.bg[0], sc.bg[1], sc.bg[2], sc.bg[3]), ctx->stream);
        ctx->n_launches += 1;
    }
    int rc = RGPU_OK;
    if (with_jobs) rc = submit(ctx, jobs, n_jobs, RGPU_BATCH_ORDERED, close_flag);
    if (rc == RGPU_OK && sc.rgba) {
        launch_to_rgba8(sc.layer, sc.rgba, n, ctx->stream);
        ctx->n_launches += 1;
    }
    return rc;
}

// Totals of a job list after build_tables() has turned it into the device form (ctx->h_jobs / ctx->h_paints).
struct Tables {
    int variant = 0;
    TileShape ts{};
    uint32_t item_acc = 0, band_acc = 0, tile_acc = 0, n_paints = 0, n_live = 0;
    uint64_t est_lines = 0;
    bool all_small = true;
    bool gradients = false;  // some job has a gradient paint
};

// rgpu_job list -> JobDev / PaintDev tables in the context's pinned staging.  Jobs that draw nothing are dropped
// (empty windows, silent no-op fills: src/rasterize.rs:87-92, 320-322).
int build_tables(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, int close_flag, SceneArgs* scene, Tables& tb) {
    int rc;
    if ((rc = ensure_pinned(ctx, ctx->h_jobs, ctx->h_jobs_cap, n_jobs))) return rc;
    if ((rc = ensure_pinned(ctx, ctx->h_paints, ctx->h_paints_cap, n_jobs))) return rc;
    // the previous batch's table upload reads these pinned buffers asynchronously: let it finish before overwriting
    CK(ctx, cudaEventSynchronize(ctx->h_tables_ev));

    // tile variant: small canvases get a (128 x 64) tile, everything else (1024 x 8)
    uint32_t max_w = 0;
    for (size_t j = 0; j < n_jobs; j++) max_w = std::max(max_w, jobs[j].width);
    const int variant = (max_w <= 128) ? 1 : 0;
    const TileShape ts = scene ? scene_tile_shape() : raster_tile_shape(variant);
    tb.variant = variant;
    tb.ts = ts;

    uint32_t item_acc = 0, band_acc = 0, tile_acc = 0, n_paints = 0;
    uint64_t est_lines = 0;
    const rgpu_paint* last_paint = nullptr;
    int last_paint_index = -1;
    const rgpu_job* last_paint_job = nullptr;
    uint32_t n_live = 0;
    bool all_small = true;
    for (size_t j = 0; j < n_jobs; j++) {
        const rgpu_job& in = jobs[j];
        if (!in.path) return fail(ctx, RGPU_ERR_INVALID, "job.path is NULL");
        if (in.width == 0 || in.height == 0) continue;  // reference: empty iterator / nothing to write
        if (!in.canvas) return fail(ctx, RGPU_ERR_INVALID, "job.canvas is NULL");
        if (in.mode == RGPU_JOB_MASK && in.width < 1) continue;
        JobDev d;
        std::memset(&d, 0, sizeof(d));
        d.paint_index = -1;
        if (in.mode < RGPU_JOB_MASK || in.mode > RGPU_JOB_RENDER) return fail(ctx, RGPU_ERR_INVALID, "unknown job mode");
        if (in.mode == RGPU_JOB_FILL || in.mode == RGPU_JOB_RENDER) {
            if (!in.paint) return fail(ctx, RGPU_ERR_INVALID, "fill job without a paint");
            if (in.paint->n_stops > RGPU_MAX_STOPS) return fail(ctx, RGPU_ERR_INVALID, "too many gradient stops (RGPU_MAX_STOPS)");
            // a solid paint is its colour: consecutive jobs with the same colour share one table entry
            bool same = last_paint_job && in.paint->kind == RGPU_PAINT_SOLID && last_paint->kind == RGPU_PAINT_SOLID &&
                        (last_paint == in.paint || std::memcmp(last_paint->solid, in.paint->solid, sizeof(in.paint->solid)) == 0);
            if (same) {
                d.paint_index = last_paint_index;
            } else {
                if (!build_paint(in, ctx->h_paints[n_paints])) {  // silent no-op fill; a RENDER job still creates its window
                    if (in.mode == RGPU_JOB_RENDER)
                        CK(ctx, cudaMemset2DAsync(static_cast<float4*>(in.canvas) + in.origin, in.row_stride * 16, 0, (size_t)in.width * 16, in.height, ctx->stream));
                    continue;
                }
                d.paint_index = (int)n_paints;
                last_paint = in.paint;
                last_paint_index = d.paint_index;
                last_paint_job = &in;
                n_paints++;
            }
        }
        std::memcpy(d.tr, in.tr, sizeof(d.tr));
        d.pts = in.path->pts;
        d.items = in.path->items;
        d.items_packed = in.path->items_packed;
        d.n_curves = in.path->n_curves;
        d.item_begin = item_acc;
        d.n_items = in.path->n_items;
        d.width_out = (int32_t)in.width;
        d.height = (int32_t)in.height;
        // reference `width`: img.width - 1 (mask: the image's own last column; mask_iter: (w+1) - 1)
        d.clamp_w = (in.mode == RGPU_JOB_MASK) ? (double)in.width - 1.0 : (double)in.width;
        d.rule = in.fill_rule;
        d.mode = in.mode;
        d.close = close_flag;
        d.fix_shift = ctx->fix_shift;
        d.y_org = ctx->job_row_origin ? ctx->job_row_origin[j] : 0.0;
        d.canvas = in.canvas;
        d.origin = in.origin;
        d.row_stride = in.row_stride;
        if (scene) {
            // the job's window inside the layer: its tile grid is the layer's (see scene.cu)
            if (in.mode != RGPU_JOB_FILL) return fail(ctx, RGPU_ERR_INVALID, "scene batches take FILL jobs only");
            if (in.canvas != (void*)scene->layer || in.row_stride != scene->width) return fail(ctx, RGPU_ERR_INVALID, "scene job is not a window of the layer");
            const size_t vx = in.origin % scene->width, vy = in.origin / scene->width;
            if (vx + in.width > scene->width || vy + in.height > scene->height) return fail(ctx, RGPU_ERR_INVALID, "scene job window leaves the layer");
            d.ox = (int32_t)(vx % ts.cw);
            d.oy = (int32_t)(vy % ts.th);
            d.sc0 = (int32_t)(vx / ts.cw);
            d.sb0 = (int32_t)(vy / ts.th);
        }
        d.band_begin = band_acc;
        d.n_bands = (in.height + d.oy + ts.th - 1) / ts.th;
        d.n_chunks = (in.width + d.ox + ts.cw - 1) / ts.cw;
        d.tile_begin = tile_acc;
        item_acc += d.n_items;
        band_acc += d.n_bands;
        tile_acc += d.n_bands * d.n_chunks;
        est_lines += (uint64_t)in.path->n_curves * 24 + (in.path->n_items - in.path->n_curves) + 16;
        all_small = all_small && !scene && d.y_org == 0.0 && small_canvas_eligible(in.width, in.height, in.mode);  // (the fused small-canvas kernel knows no row origin)
        ctx->h_jobs[n_live++] = d;
    }
    tb.item_acc = item_acc;
    tb.band_acc = band_acc;
    tb.tile_acc = tile_acc;
    tb.n_paints = n_paints;
    tb.n_live = n_live;
    tb.est_lines = est_lines;
    tb.all_small = all_small;
    for (uint32_t k = 0; k < n_paints; k++) tb.gradients = tb.gradients || ctx->h_paints[k].kind != 0;
    return RGPU_OK;
}

int submit(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags, int close_flag, bool ordered_lines, SceneArgs* scene) {
    if (n_jobs == 0 && !scene) return RGPU_OK;
    if (!jobs && n_jobs) return fail(ctx, RGPU_ERR_INVALID, "jobs is NULL");
    if (scene && (ctx->two_pass || n_jobs == 0)) return scene_fallback(ctx, jobs, n_jobs, close_flag, *scene, n_jobs != 0);
    if (!(ctx->flatness > 0.0)) return fail(ctx, RGPU_ERR_INVALID, "flatness must be > 0 (the reference loops forever on 0)");
    int rc;
    Tables tb;
    if ((rc = build_tables(ctx, jobs, n_jobs, close_flag, scene, tb))) return rc;
    const int variant = tb.variant;
    const TileShape ts = tb.ts;
    const uint32_t item_acc = tb.item_acc, tile_acc = tb.tile_acc, n_paints = tb.n_paints, n_live = tb.n_live;
    const uint64_t est_lines = tb.est_lines;
    const bool all_small = tb.all_small;
    ctx->need_lines = ctx->need_refs = 0;
    if (n_live == 0) {
        std::memset(ctx->h_status, 0, sizeof(Status));
        ctx->d_status_cur = nullptr;
        if (scene) return scene_fallback(ctx, jobs, n_jobs, close_flag, *scene, false);
        return RGPU_OK;
    }
    if (!(all_small && !ordered_lines)) {
        // RENDER on the tiled path: clear the window, then an ordinary FILL (the fused small-canvas kernel writes it once)
        for (uint32_t k = 0; k < n_live; k++) {
            JobDev& d = ctx->h_jobs[k];
            if (d.mode != kModeRender) continue;
            CK(ctx, cudaMemset2DAsync(static_cast<float4*>(d.canvas) + d.origin, d.row_stride * 16, 0, (size_t)d.width_out * 16, d.height, ctx->stream));
            d.mode = kModeFill;
        }
    }
    if (all_small && !ordered_lines) {
        // every canvas fits one CTA's shared memory: a single fused kernel per launch, nothing else touches HBM
        if ((rc = ensure_dev(ctx, ctx->jobs, sizeof(JobDev) * n_live))) return rc;
        if ((rc = ensure_dev(ctx, ctx->paints, sizeof(PaintDev) * std::max<uint32_t>(n_paints, 1)))) return rc;
        cudaStream_t s = ctx->stream;
        JobDev* d_jobs = static_cast<JobDev*>(ctx->jobs.p);
        PaintDev* d_paints = static_cast<PaintDev*>(ctx->paints.p);
        Status* d_status = static_cast<Status*>(ctx->status.p);
        bool uploaded = false;
        if ((rc = upload_table(ctx, ctx->jobs_shadow, d_jobs, ctx->h_jobs, sizeof(JobDev) * n_live, uploaded))) return rc;
        if (n_paints && (rc = upload_table(ctx, ctx->paints_shadow, d_paints, ctx->h_paints, sizeof(PaintDev) * n_paints, uploaded))) return rc;
        if (uploaded) CK(ctx, cudaEventRecord(ctx->h_tables_ev, s));
        CK(ctx, cudaMemsetAsync(d_status, 0, sizeof(Status), s));
        const double thr = 16.0 * ctx->flatness * ctx->flatness;  // PathFlattenIter::new, src/path.rs:749
        const bool prof = ctx->profiling;
        ctx->ev_valid = false;
        if (prof) {
            CK(ctx, cudaEventRecord(ctx->ev[0], s));
            CK(ctx, cudaEventRecord(ctx->ev[1], s));
            CK(ctx, cudaEventRecord(ctx->ev[2], s));
        }
        if (flags & RGPU_BATCH_INDEPENDENT) {
            launch_small_canvas(d_jobs, 0, n_live, d_paints, thr, d_status, tb.gradients, s);
            ctx->n_launches += 1;
        } else {
            for (uint32_t j = 0; j < n_live; j++) launch_small_canvas(d_jobs, j, 1, d_paints, thr, d_status, tb.gradients, s);
            ctx->n_launches += n_live;
        }
        if (prof) {
            CK(ctx, cudaEventRecord(ctx->ev[3], s));
            ctx->ev_valid = true;
        }
        ctx->d_status_cur = d_status;
        CK(ctx, cudaGetLastError());
        return RGPU_OK;
    }
    uint64_t total_slots64 = (uint64_t)item_acc * kSlotsPerItem + 1;
    if (total_slots64 > 0x7fffffffull) return fail(ctx, RGPU_ERR_INVALID, "batch too large (more than 2^28 path items)");
    uint32_t total_slots = item_acc * kSlotsPerItem;
    ctx->last_total_slots = total_slots;
    if ((rc = ensure_dev(ctx, ctx->jobs, sizeof(JobDev) * n_live))) return rc;
    if ((rc = ensure_dev(ctx, ctx->paints, sizeof(PaintDev) * std::max<uint32_t>(n_paints, 1)))) return rc;
    cudaStream_t s = ctx->stream;
    JobDev* d_jobs = static_cast<JobDev*>(ctx->jobs.p);
    PaintDev* d_paints = static_cast<PaintDev*>(ctx->paints.p);
    const double thr = 16.0 * ctx->flatness * ctx->flatness;  // PathFlattenIter::new, src/path.rs:749
    const bool prof = ctx->profiling;
    ctx->ev_valid = false;

    if (ordered_lines) {
        // `Path::flatten` parity: count -> scan -> emit in the reference's order; nothing is rasterized
        size_t want_lines = std::max<uint64_t>(ctx->lines_cap, est_lines);
        if ((rc = ensure_dev(ctx, ctx->slot_counts, sizeof(uint32_t) * (total_slots + 1)))) return rc;
        if ((rc = ensure_dev(ctx, ctx->slot_offs, sizeof(uint32_t) * (total_slots + 1)))) return rc;
        if ((rc = ensure_dev(ctx, ctx->lines, sizeof(double4) * want_lines))) return rc;
        ctx->lines_cap = std::min<size_t>(ctx->lines.cap / sizeof(double4), 0xfffffff0u);
        if ((rc = ensure_dev(ctx, ctx->scan_temp, scan_temp_bytes(total_slots + 1)))) return rc;
        Status* d_status = static_cast<Status*>(ctx->status.p);
        uint32_t* d_counts = static_cast<uint32_t*>(ctx->slot_counts.p);
        uint32_t* d_offs = static_cast<uint32_t*>(ctx->slot_offs.p);
        bool uploaded = false;
        if ((rc = upload_table(ctx, ctx->jobs_shadow, d_jobs, ctx->h_jobs, sizeof(JobDev) * n_live, uploaded))) return rc;
        if (uploaded) CK(ctx, cudaEventRecord(ctx->h_tables_ev, s));
        CK(ctx, cudaMemsetAsync(d_status, 0, sizeof(Status), s));
        launch_flatten_count(d_jobs, n_live, item_acc, thr, d_counts, d_status, s);
        launch_exclusive_scan(d_counts, d_offs, total_slots + 1, ctx->scan_temp.p, ctx->scan_temp.cap, s);
        launch_flatten_emit(d_jobs, n_live, item_acc, thr, d_offs, static_cast<double4*>(ctx->lines.p), (uint32_t)ctx->lines_cap, d_status, s);
        CK(ctx, cudaMemcpyAsync(&d_status->n_lines, d_offs + total_slots, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        ctx->n_launches += 3;
        ctx->d_status_cur = d_status;
        CK(ctx, cudaGetLastError());
        return RGPU_OK;
    }

    // packed flatten grid: per job (n_curves << depth) curve-slot threads, then one thread per line item, padded to whole warps
    const int cut_depth = flatten_cut_depth(item_acc);
    uint32_t thread_acc = 0;
    for (uint32_t j = 0; j < n_live; j++) {
        JobDev& d = ctx->h_jobs[j];
        d.thread_begin = thread_acc;
        const uint64_t n = ((uint64_t)d.n_curves << cut_depth) + (d.n_items - d.n_curves);
        if (thread_acc + ((n + 31) & ~31ull) > 0x7fffffffull) return fail(ctx, RGPU_ERR_INVALID, "batch too large");
        thread_acc += (uint32_t)((n + 31) & ~31ull);
    }

    // ---- raster path -------------------------------------------------------------------------------------------
    // fixed bins (default): [flatten + write fixed-capacity bins] -> raster               (2 launches)
    // two-pass (fallback):  [flatten + count per tile] -> scan -> [flatten + write bins] -> raster
    uint32_t bin_cap = 0;
    if (!ctx->two_pass) {
        // first guess: four times the average load the line estimate predicts; a tile that wants more raises
        // refs_overflow + bin_max and the *_sync entry points re-run with the exact capacity
        uint64_t avg = (est_lines + est_lines / 2) / std::max<uint32_t>(tile_acc, 1) + 1;
        uint32_t guess = 64;
        while (guess < 4 * avg && guess < 4096) guess <<= 1;
        bin_cap = std::max(ctx->bin_cap, guess);
        if ((uint64_t)bin_cap * tile_acc * sizeof(double4) > kFixedBinBudget) bin_cap = 0;
    }
    const bool fixed = bin_cap != 0;
    if (scene && !fixed) return scene_fallback(ctx, jobs, n_jobs, close_flag, *scene, true);
    ctx->last_fixed = fixed;
    ctx->last_tiles = tile_acc;
    size_t want_refs = fixed ? (size_t)bin_cap * tile_acc : std::max<uint64_t>(ctx->refs_cap, est_lines + est_lines / 2);
    if ((rc = ensure_dev(ctx, ctx->refs, sizeof(double4) * want_refs))) return rc;
    ctx->refs_cap = std::min<size_t>(ctx->refs.cap / sizeof(double4), 0xfffffff0u);
    const uint32_t n_raster_launches = ((flags & RGPU_BATCH_INDEPENDENT) || scene) ? 1u : n_live;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    // carry look-back state, validated by epoch (cleared only when (re)allocated or when the epoch wraps)
    {
        size_t need = sizeof(unsigned long long) * kStateRows * (size_t)tile_acc;
        size_t before = ctx->tile_state.cap;
        if ((rc = ensure_dev(ctx, ctx->tile_state, need))) return rc;
        ctx->epoch++;
        if (ctx->tile_state.cap != before || ctx->epoch >= (1u << 30)) {
            CK(ctx, cudaMemsetAsync(ctx->tile_state.p, 0, ctx->tile_state.cap, ctx->stream));
            if (ctx->epoch >= (1u << 30)) ctx->epoch = 1;
        }
    }
    Status* d_status = nullptr;
    Status* d_status_next = nullptr;
    uint32_t *d_tickets = nullptr, *d_bc = nullptr, *d_cur = nullptr, *d_bo = nullptr;
    void* d_scan = nullptr;
    if (fixed) {
        // Self-cleaning block [status A | status B | raster tickets | tile counters]: the kernels leave the tickets and
        // counters zero (raster.cu) and clear the OTHER status block for the batch after this one (flatten.cu), so the
        // steady state needs no memset.  Only growing the block clears it wholesale.
        if (n_raster_launches > ctx->fx_tickets_cap || tile_acc > ctx->fx_tiles_cap) {
            const size_t tcap = std::max<size_t>(n_raster_launches + n_raster_launches / 2, 64);
            const size_t ccap = std::max<size_t>((size_t)tile_acc + tile_acc / 2, 1024);
            const size_t bytes = 512 + up(tcap * 4) + up(ccap * 4);
            if ((rc = ensure_dev(ctx, ctx->fixed_block, bytes))) return rc;
            CK(ctx, cudaMemsetAsync(ctx->fixed_block.p, 0, ctx->fixed_block.cap, s));
            ctx->fx_tickets_cap = (uint32_t)tcap;
            ctx->fx_tiles_cap = (uint32_t)ccap;
        }
        char* fb = static_cast<char*>(ctx->fixed_block.p);
        d_status = reinterpret_cast<Status*>(fb + (ctx->fx_parity ? 256 : 0));
        d_status_next = reinterpret_cast<Status*>(fb + (ctx->fx_parity ? 0 : 256));
        // fx_parity is toggled only once the flatten launch (which clears the other block) is issued: an early return
        // on the fallible steps below must not leave the next batch on a block nobody cleared
        d_tickets = reinterpret_cast<uint32_t*>(fb + 512);
        d_bc = reinterpret_cast<uint32_t*>(fb + 512 + up((size_t)ctx->fx_tickets_cap * 4));
        d_bo = d_bc;
        if (item_acc == 0) CK(ctx, cudaMemsetAsync(d_status_next, 0, sizeof(Status), s));  // no flatten launch to do it
    } else {
        // two-pass: one block that must be zero at the start of the batch, cleared by ONE memset:
        // [status | raster tickets | tile_counts | tile_cursor | scan tile states]
        const size_t tickets_off = up(sizeof(Status));
        const size_t counts_off = tickets_off + up((size_t)n_raster_launches * 4);
        const size_t cursor_off = counts_off + up((size_t)(tile_acc + 1) * 4);
        const size_t scan_off = cursor_off + up((size_t)(tile_acc + 1) * 4);
        const size_t zero_bytes = scan_off + up(scan_temp_bytes(tile_acc + 1));
        if ((rc = ensure_dev(ctx, ctx->zero_block, zero_bytes))) return rc;
        if ((rc = ensure_dev(ctx, ctx->tile_offs, sizeof(uint32_t) * (tile_acc + 1)))) return rc;
        char* zb = static_cast<char*>(ctx->zero_block.p);
        d_status = reinterpret_cast<Status*>(zb);
        d_tickets = reinterpret_cast<uint32_t*>(zb + tickets_off);
        d_bc = reinterpret_cast<uint32_t*>(zb + counts_off);
        d_cur = reinterpret_cast<uint32_t*>(zb + cursor_off);
        d_scan = zb + scan_off;
        d_bo = static_cast<uint32_t*>(ctx->tile_offs.p);
        CK(ctx, cudaMemsetAsync(zb, 0, zero_bytes, s));
    }
    double4* d_refs = static_cast<double4*>(ctx->refs.p);
    unsigned long long* d_state = static_cast<unsigned long long*>(ctx->tile_state.p);

    if (scene) {
        // per-band job lists (CSR, submission order inside a band): [offsets: n_bands + 1 | job indices]
        scene->n_bands = (scene->height + ts.th - 1) / ts.th;
        scene->n_chunks = (scene->width + ts.cw - 1) / ts.cw;
        std::vector<uint32_t>& L = ctx->h_scene_lists;
        const uint32_t nb = scene->n_bands;
        L.assign(nb + 1, 0u);
        for (uint32_t j = 0; j < n_live; j++) {
            const JobDev& d = ctx->h_jobs[j];
            for (uint32_t b = 0; b < d.n_bands; b++) L[d.sb0 + b + 1]++;
        }
        for (uint32_t b = 0; b < nb; b++) L[b + 1] += L[b];
        const uint32_t total = L[nb];
        L.resize(nb + 1 + total);
        std::vector<uint32_t> cur(L.begin(), L.begin() + nb);
        for (uint32_t j = 0; j < n_live; j++) {
            const JobDev& d = ctx->h_jobs[j];
            for (uint32_t b = 0; b < d.n_bands; b++) L[nb + 1 + cur[d.sb0 + b]++] = j;
        }
        {
            // band order: estimated work (sum of the window widths of the band's jobs) descending, ties by index
            std::vector<std::pair<uint64_t, uint32_t>> key(nb);
            for (uint32_t b = 0; b < nb; b++) {
                uint64_t wsum = 0;
                for (uint32_t k = L[b]; k < L[b + 1]; k++) wsum += (uint64_t)ctx->h_jobs[L[nb + 1 + k]].width_out;
                key[b] = {~wsum, b};  // ascending sort of (~work, index)
            }
            std::sort(key.begin(), key.end());
            const size_t base = L.size();
            L.resize(base + nb);
            for (uint32_t b = 0; b < nb; b++) L[base + b] = key[b].second;
        }
        if ((rc = ensure_dev(ctx, ctx->scene_lists, sizeof(uint32_t) * L.size()))) return rc;
        bool up_lists = false;
        if ((rc = upload_table(ctx, ctx->lists_shadow, ctx->scene_lists.p, L.data(), sizeof(uint32_t) * L.size(), up_lists))) return rc;
        scene->band_offs = static_cast<const uint32_t*>(ctx->scene_lists.p);
        scene->band_jobs = scene->band_offs + nb + 1;
        scene->band_order = scene->band_offs + (L.size() - nb);
    }
    // a single job travels in the kernel parameters; a table is uploaded only for multi-job batches
    {
        bool uploaded = false;
        if ((n_live > 1 || !fixed || scene) && (rc = upload_table(ctx, ctx->jobs_shadow, d_jobs, ctx->h_jobs, sizeof(JobDev) * n_live, uploaded))) return rc;
        if (n_paints && (rc = upload_table(ctx, ctx->paints_shadow, d_paints, ctx->h_paints, sizeof(PaintDev) * n_paints, uploaded))) return rc;
        if (uploaded) CK(ctx, cudaEventRecord(ctx->h_tables_ev, s));
    }
    if (prof) CK(ctx, cudaEventRecord(ctx->ev[0], s));
    if (fixed) {
        ctx->fx_parity ^= 1u;
        launch_flatten_bin_fixed(d_jobs, ctx->h_jobs, n_live, thread_acc, cut_depth, thr, d_bc, d_refs, bin_cap, ts.th, ts.cw, d_status, d_status_next, s);
        ctx->n_launches += 1;
        if (prof) CK(ctx, cudaEventRecord(ctx->ev[1], s));
    } else {
        launch_flatten_bin_count(d_jobs, n_live, item_acc, thr, d_bc, ts.th, ts.cw, d_status, s);
        if (prof) CK(ctx, cudaEventRecord(ctx->ev[1], s));
        launch_exclusive_scan(d_bc, d_bo, tile_acc + 1, d_scan, 0, s, /*temp_is_zero=*/true);
        launch_flatten_bin_emit(d_jobs, n_live, item_acc, thr, d_bo, tile_acc, d_cur, d_refs, (uint32_t)ctx->refs_cap, ts.th, ts.cw, d_status, s);
        ctx->n_launches += 3;
    }
    if (prof) CK(ctx, cudaEventRecord(ctx->ev[2], s));
    static const bool no_pdl = getenv("RGPU_NO_PDL") != nullptr;  // A/B switch
    const bool pdl_ok = fixed && !prof && !no_pdl;
    const bool zero_early = est_lines + est_lines / 2 >= 2ull * tile_acc;  // most tiles will hold lines
    if (scene) {
        launch_scene(d_jobs, n_live, d_paints, d_bo, bin_cap, d_refs, d_state, ctx->epoch, d_tickets, d_status, *scene, /*pdl=*/pdl_ok && item_acc != 0, s);
        ctx->n_launches += 1;
    } else if (flags & RGPU_BATCH_INDEPENDENT) {
        launch_raster(variant, d_jobs, ctx->h_jobs, n_live, 0, 0, tile_acc, d_paints, d_bo, bin_cap, d_refs, d_state, ctx->epoch, d_tickets, d_status, zero_early, /*pdl=*/pdl_ok ? 1 : 0, s);
        ctx->n_launches += 1;
    } else {
        for (uint32_t j = 0; j < n_live; j++) {
            const JobDev& d = ctx->h_jobs[j];
            // a FILL launch spends its time evaluating the paint per pixel: same tile, four times the threads
            launch_raster((variant == 0 && d.mode == kModeFill && d.paint_index >= 0 && ctx->h_paints[d.paint_index].kind != 0) ? 2 : variant,
                          d_jobs, ctx->h_jobs, 1, j, d.tile_begin, d.n_bands * d.n_chunks, d_paints, d_bo, bin_cap, d_refs, d_state,
                          ctx->epoch, d_tickets + j, d_status, zero_early, /*pdl=*/pdl_ok ? (j == 0 ? 1 : 2) : 0, s);
            ctx->n_launches += 1;
        }
    }
    if (prof) {
        CK(ctx, cudaEventRecord(ctx->ev[3], s));
        ctx->ev_valid = true;
    }
    ctx->d_status_cur = d_status;
    CK(ctx, cudaGetLastError());
    return RGPU_OK;
}

int check_status(rgpu_ctx* ctx) {
    // the device-side status of the last submission is fetched only here (keeps the submission path copy-free)
    if (ctx->d_status_cur) {
        CK(ctx, cudaMemcpyAsync(ctx->h_status, ctx->d_status_cur, sizeof(Status), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d_status_cur = nullptr;
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    const Status& st = *ctx->h_status;
    if (st.nan_flag) return fail(ctx, RGPU_ERR_NAN, "cannot flatten segment with NaN");
    if (st.depth_flag) return fail(ctx, RGPU_ERR_DEPTH, "curve subdivision exceeded the device stack depth");
    ctx->need_lines = st.n_lines;
    ctx->need_refs = st.n_refs;
    if (st.lines_overflow || st.refs_overflow) return fail(ctx, RGPU_ERR_CAPACITY, "internal scratch overflow (retry grows it)");
    if (st.winding_flag)
        return fail(ctx, RGPU_ERR_WINDING, "non-zero winding number beyond the guard of the 32-bit cells (the *_sync entry points re-run such a batch with "
                                           "a wider integer part; rgpu_set_winding_bits selects it up front)");
    ctx->last_lines = st.n_lines;
    ctx->last_refs = st.n_refs;
    return RGPU_OK;
}

int submit_sync(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags, int close_flag, bool ordered_lines = false,
                SceneArgs* scene = nullptr) {
    for (int attempt = 0; attempt < 4; attempt++) {
        int rc = submit(ctx, jobs, n_jobs, flags, close_flag, ordered_lines, scene);
        if (rc) return rc;
        rc = check_status(ctx);
        if (rc == RGPU_ERR_WINDING && ctx->fix_shift > kFixShiftWide) {
            // a NonZero winding reached the guard of the Q7.24 cells: once more in Q13.18 (windings up to +-8192).  Only batches
            // that overwrite their output can be repeated: a FILL has already blended into its canvas (the caller re-renders
            // with rgpu_set_winding_bits(ctx, 14); rgpu_fill restores the image from the host copy and repeats itself).
            bool repeatable = !(scene && !scene->fresh);
            for (size_t j = 0; j < n_jobs && repeatable; j++) repeatable = scene || jobs[j].mode != RGPU_JOB_FILL;
            if (!repeatable) return rc;
            const int keep = ctx->fix_shift;
            ctx->fix_shift = kFixShiftWide;
            rc = submit(ctx, jobs, n_jobs, flags, close_flag, ordered_lines, scene);
            if (rc == RGPU_OK) rc = check_status(ctx);
            ctx->fix_shift = keep;
            if (rc != RGPU_ERR_CAPACITY) return rc;
        }
        if (rc != RGPU_ERR_CAPACITY) return rc;
        // grow: line count is exact once the emit pass overflowed (it comes from the scan); the reference
        // count is only known when the lines fitted, so over-provision it from the line count.
        const Status& st = *ctx->h_status;
        size_t nl = std::max<size_t>(st.n_lines, ctx->lines_cap);
        if (st.lines_overflow) {
            uint32_t total = st.n_lines;  // fused flatten: every CTA added its exact count before checking capacity
            if (ordered_lines) {
                // n_lines is written by bin_count, which is skipped on overflow: read the scan total instead
                uint32_t* d_offs = static_cast<uint32_t*>(ctx->slot_offs.p);
                CK(ctx, cudaMemcpy(&total, d_offs + ctx->last_total_slots, sizeof(uint32_t), cudaMemcpyDeviceToHost));
            }
            nl = (size_t)total + total / 16 + 64;
        }
        int rc2;
        if (ordered_lines) {
            if ((rc2 = ensure_dev(ctx, ctx->lines, sizeof(double4) * nl))) return rc2;
            ctx->lines_cap = ctx->lines.cap / sizeof(double4);
        } else if (ctx->last_fixed) {
            // a tile wanted more lines than its fixed bin holds: bin_max is exact (every line took a slot number),
            // so one re-run with that capacity succeeds
            uint32_t want = st.bin_max + st.bin_max / 8 + 8;
            want = (want + 31u) & ~31u;
            ctx->bin_cap = std::max(ctx->bin_cap, want);  // submit() falls back to two-pass if this blows the budget
        } else {
            size_t nr = st.refs_overflow ? (size_t)st.n_refs + st.n_refs / 16 + 64 : std::max<size_t>(ctx->refs_cap, nl * 2);
            if ((rc2 = ensure_dev(ctx, ctx->refs, sizeof(double4) * nr))) return rc2;
            ctx->refs_cap = ctx->refs.cap / sizeof(double4);
        }
    }
    return fail(ctx, RGPU_ERR_CAPACITY, "internal scratch overflow persisted after 4 attempts");
}

}  // namespace

extern "C" {

int rgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* rgpu_name(void) { return "gpu-signed-difference"; }

int rgpu_create(int device, double flatness, rgpu_ctx** out) {
    if (!out) return RGPU_ERR_INVALID;
    *out = nullptr;
    if (!(flatness > 0.0)) {
        g_create_err = "flatness must be > 0";
        return RGPU_ERR_INVALID;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e) + " (rasterize_b200 has no CPU fallback)";
        return RGPU_ERR_CUDA;
    }
    auto* ctx = new rgpu_ctx();
    ctx->device = device;
    ctx->flatness = flatness;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_status), sizeof(Status));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->status.p, 256);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->h_tables_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        g_create_err = std::string("rgpu_create: ") + cudaGetErrorString(e);
        delete ctx;
        return RGPU_ERR_CUDA;
    }
    ctx->status.cap = 256;
    std::memset(ctx->h_status, 0, sizeof(Status));
    {
        // The device paths rgpu_path_stroke / rgpu_parse_svg_batch hand out come from the device's stream-ordered pool
        // (cudaMallocAsync): keep what is freed in the pool instead of returning it to the driver at every synchronisation,
        // so that a loop of stroke / parse calls does not pay cudaMalloc each time.
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = ctx;
    return RGPU_OK;
}

void rgpu_destroy(rgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->jobs, &ctx->paints, &ctx->slot_counts, &ctx->slot_offs, &ctx->lines, &ctx->line_job, &ctx->zero_block, &ctx->tile_offs,
                      &ctx->tile_state, &ctx->fixed_block, &ctx->refs, &ctx->scan_temp, &ctx->status, &ctx->img_f32, &ctx->img_f64, &ctx->img_lin, &ctx->tmp_pts, &ctx->tmp_items, &ctx->scene_lists, &ctx->stroke_buf, &ctx->px_counts, &ctx->px_out, &ctx->rc_buf, &ctx->rc_lits};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < rgpu_ctx::kRing; i++) {
        if (ctx->ring_slab[i].p) cudaFree(ctx->ring_slab[i].p);
        if (ctx->ring_rgba[i].p) cudaFree(ctx->ring_rgba[i].p);
        if (ctx->h_alpha[i]) cudaFreeHost(ctx->h_alpha[i]);
        if (ctx->ring_done[i]) cudaEventDestroy(ctx->ring_done[i]);
        if (ctx->ring_copied[i]) cudaEventDestroy(ctx->ring_copied[i]);
        for (cudaEvent_t e : ctx->ring_alpha[i]) cudaEventDestroy(e);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_jobs) cudaFreeHost(ctx->h_jobs);
    if (ctx->h_paints) cudaFreeHost(ctx->h_paints);
    if (ctx->h_items) cudaFreeHost(ctx->h_items);
    if (ctx->h_rc) cudaFreeHost(ctx->h_rc);
    if (ctx->h_lits) cudaFreeHost(ctx->h_lits);
    for (int i = 0; i < 2; i++)
        if (ctx->h_chunk[i]) cudaFreeHost(ctx->h_chunk[i]);
    if (ctx->h_pts) cudaFreeHost(ctx->h_pts);
    for (int i = 0; i < 4; i++)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (cudaEvent_t e : ctx->chunk_ev) cudaEventDestroy(e);
    if (ctx->h_tables_ev) cudaEventDestroy(ctx->h_tables_ev);
    ctx->pool.reset();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* rgpu_last_error(const rgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

void* rgpu_stream(rgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int rgpu_sync(rgpu_ctx* ctx) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RGPU_OK;
}

int rgpu_device_alloc(rgpu_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return RGPU_OK;
}
int rgpu_device_free(rgpu_ctx* ctx, void* p) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaFree(p));
    return RGPU_OK;
}
int rgpu_device_zero(rgpu_ctx* ctx, void* p, size_t bytes) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaMemsetAsync(p, 0, bytes, ctx->stream));
    return RGPU_OK;
}
int rgpu_memcpy_h2d(rgpu_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RGPU_OK;
}
int rgpu_memcpy_d2h(rgpu_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RGPU_OK;
}
int rgpu_host_alloc(rgpu_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMallocHost(out, bytes ? bytes : 1));
    return RGPU_OK;
}
int rgpu_host_free(rgpu_ctx* ctx, void* p) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaFreeHost(p));
    return RGPU_OK;
}

int rgpu_path_upload(rgpu_ctx* ctx, const rgpu_path* path, rgpu_dpath** out) {
    if (!ctx || !out) return RGPU_ERR_INVALID;
    *out = nullptr;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = validate_path(ctx, path);
    if (rc) return rc;
    auto* dp = new rgpu_dpath();
    rc = upload_path(ctx, path, dp);
    if (rc) {
        free_path(dp);
        delete dp;
        return rc;
    }
    *out = dp;
    return RGPU_OK;
}

void rgpu_path_free(rgpu_ctx* ctx, rgpu_dpath* p) {
    if (!p) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    free_path(p);
    delete p;
}

int rgpu_render_batch(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    return submit(ctx, jobs, n_jobs, flags, 1);
}

int rgpu_set_winding_bits(rgpu_ctx* ctx, int integer_bits) {
    if (!ctx) return RGPU_ERR_INVALID;
    if (integer_bits != 8 && integer_bits != 14) return fail(ctx, RGPU_ERR_INVALID, "winding bits: 8 (Q7.24 cells, the default) or 14 (Q13.18)");
    ctx->fix_shift = integer_bits == 8 ? kFixShift : kFixShiftWide;
    return RGPU_OK;
}

int rgpu_batch_status(rgpu_ctx* ctx) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    return check_status(ctx);
}

int rgpu_render_batch_sync(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, uint32_t flags) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    return submit_sync(ctx, jobs, n_jobs, flags, 1);
}

static int scene_args(rgpu_ctx* ctx, float* layer_dev, size_t width, size_t height, int fresh, const float* bg, uint8_t* rgba_dev, SceneArgs& sc) {
    if (!layer_dev) return fail(ctx, RGPU_ERR_INVALID, "layer is NULL");
    if (width > 0x3fffffffu || height > 0x3fffffffu) return fail(ctx, RGPU_ERR_INVALID, "layer too large");
    std::memset(&sc, 0, sizeof(sc));
    sc.layer = reinterpret_cast<float4*>(layer_dev);
    sc.rgba = reinterpret_cast<uchar4*>(rgba_dev);
    sc.width = (uint32_t)width;
    sc.height = (uint32_t)height;
    sc.fresh = fresh != 0;
    sc.store_lin = 1;
    if (fresh && bg) std::memcpy(sc.bg, bg, sizeof(sc.bg));
    return RGPU_OK;
}

int rgpu_render_scene(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, float* layer_dev, size_t width, size_t height, int fresh,
                      const float* bg, uint8_t* rgba_dev) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (width == 0 || height == 0) return RGPU_OK;
    SceneArgs sc;
    int rc = scene_args(ctx, layer_dev, width, height, fresh, bg, rgba_dev, sc);
    if (rc) return rc;
    return submit(ctx, jobs, n_jobs, RGPU_BATCH_ORDERED, 1, false, &sc);
}

int rgpu_render_scene_sync(rgpu_ctx* ctx, const rgpu_job* jobs, size_t n_jobs, float* layer_dev, size_t width, size_t height, int fresh,
                           const float* bg, uint8_t* rgba_dev) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (width == 0 || height == 0) return RGPU_OK;
    SceneArgs sc;
    int rc = scene_args(ctx, layer_dev, width, height, fresh, bg, rgba_dev, sc);
    if (rc) return rc;
    return submit_sync(ctx, jobs, n_jobs, RGPU_BATCH_ORDERED, 1, false, &sc);
}

// `Scene::render` of a Fill-only pipeline + export with HOST buffers: all paths are staged into the context's grow-only
// scratch with two copies (points, items), the scene compositor renders the layer in one raster launch and the image comes
// back as RGBA8 (4 B per pixel) and / or LinColor.
int rgpu_render_scene_host(rgpu_ctx* ctx, const rgpu_scene_fill* fills, size_t n_fills, size_t width, size_t height, const float* bg,
                           float* lin_out, uint8_t* rgba_out) {
    if (!ctx || (!fills && n_fills)) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (width == 0 || height == 0) return RGPU_OK;
    if (!lin_out && !rgba_out) return fail(ctx, RGPU_ERR_INVALID, "no output image");
    int rc;
    size_t n_points = 0, n_items2 = 0;
    std::vector<rgpu_dpath> dps(n_fills);
    std::vector<std::vector<uint2>> items(n_fills);
    for (size_t i = 0; i < n_fills; i++) {
        const rgpu_scene_fill& f = fills[i];
        if ((rc = validate_path(ctx, f.path))) return rc;
        if (!f.paint) return fail(ctx, RGPU_ERR_INVALID, "fill without a paint");
        if (f.paint->n_stops > RGPU_MAX_STOPS) return fail(ctx, RGPU_ERR_INVALID, "too many gradient stops");
        if ((size_t)f.x + f.width > width || (size_t)f.y + f.height > height) return fail(ctx, RGPU_ERR_INVALID, "fill window leaves the layer");
        std::vector<uint2> packed;
        build_items(f.path, items[i], dps[i].n_curves);
        pack_items(items[i], packed);
        dps[i].n_points = f.path->n_points;
        dps[i].n_items = (uint32_t)items[i].size();
        items[i].insert(items[i].end(), packed.begin(), packed.end());  // [reference order | curves first]
        n_points += f.path->n_points;
        n_items2 += items[i].size();
    }
    const size_t pts_bytes = sizeof(double2) * n_points, items_bytes = sizeof(uint2) * n_items2;
    if ((rc = ensure_dev(ctx, ctx->tmp_pts, std::max<size_t>(pts_bytes, 16)))) return rc;
    if ((rc = ensure_dev(ctx, ctx->tmp_items, std::max<size_t>(items_bytes, 16)))) return rc;
    ctx->staged_items_valid = false;  // tmp_items is about to hold this scene's item lists
    if ((rc = ensure_stage(ctx, std::max<size_t>(pts_bytes + items_bytes, 16)))) return rc;
    // every host-buffer entry point ends with a stream sync: the staging buffer is free
    char* st = static_cast<char*>(ctx->h_stage);
    size_t po = 0, io = 0;
    for (size_t i = 0; i < n_fills; i++) {
        dps[i].pts = static_cast<double2*>(ctx->tmp_pts.p) + po;
        dps[i].items = static_cast<uint2*>(ctx->tmp_items.p) + io;
        dps[i].items_packed = dps[i].items + dps[i].n_items;
        if (dps[i].n_points) std::memcpy(st + sizeof(double2) * po, fills[i].path->points, sizeof(double2) * dps[i].n_points);
        if (!items[i].empty()) std::memcpy(st + pts_bytes + sizeof(uint2) * io, items[i].data(), sizeof(uint2) * items[i].size());
        po += dps[i].n_points;
        io += items[i].size();
    }
    if (pts_bytes) CK(ctx, cudaMemcpyAsync(ctx->tmp_pts.p, st, pts_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (items_bytes) CK(ctx, cudaMemcpyAsync(ctx->tmp_items.p, st + pts_bytes, items_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const size_t n_px = width * height;
    if ((rc = ensure_dev(ctx, ctx->img_lin, sizeof(float4) * n_px))) return rc;
    if (rgba_out && (rc = ensure_dev(ctx, ctx->img_f32, 4 * n_px))) return rc;  // RGBA8 staging shares the f32 mask scratch
    std::vector<rgpu_job> jobs;
    jobs.reserve(n_fills);
    for (size_t i = 0; i < n_fills; i++) {
        const rgpu_scene_fill& f = fills[i];
        if (f.path->n_segments == 0 || f.path->n_subpaths == 0 || f.width == 0 || f.height == 0) continue;
        rgpu_job j;
        std::memset(&j, 0, sizeof(j));
        j.path = &dps[i];
        std::memcpy(j.tr, f.tr, sizeof(j.tr));
        j.fill_rule = f.fill_rule;
        j.mode = RGPU_JOB_FILL;
        j.paint = f.paint;
        j.path_bbox = f.path_bbox;
        j.canvas = ctx->img_lin.p;
        j.origin = (size_t)f.y * width + f.x;
        j.row_stride = width;
        j.width = f.width;
        j.height = f.height;
        jobs.push_back(j);
    }
    SceneArgs sc;
    if ((rc = scene_args(ctx, static_cast<float*>(ctx->img_lin.p), width, height, 1, bg, rgba_out ? static_cast<uint8_t*>(ctx->img_f32.p) : nullptr, sc)))
        return rc;
    if ((rc = submit_sync(ctx, jobs.data(), jobs.size(), RGPU_BATCH_ORDERED, 1, false, &sc))) return rc;
    if (rgba_out) CK(ctx, cudaMemcpyAsync(rgba_out, ctx->img_f32.p, 4 * n_px, cudaMemcpyDeviceToHost, ctx->stream));
    if (lin_out) CK(ctx, cudaMemcpyAsync(lin_out, ctx->img_lin.p, sizeof(float4) * n_px, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->last_h2d_bytes = pts_bytes + items_bytes;
    ctx->last_d2h_bytes = (rgba_out ? 4 * n_px : 0) + (lin_out ? sizeof(float4) * n_px : 0);
    return RGPU_OK;
}

int rgpu_last_counts(rgpu_ctx* ctx, uint64_t* n_lines, uint64_t* n_line_refs, uint64_t* n_launches) {
    if (!ctx) return RGPU_ERR_INVALID;
    if (n_lines) *n_lines = ctx->last_lines;
    if (n_line_refs) *n_line_refs = ctx->last_refs;
    if (n_launches) *n_launches = ctx->n_launches;
    return RGPU_OK;
}

int rgpu_last_transfer_bytes(rgpu_ctx* ctx, uint64_t* h2d, uint64_t* d2h) {
    if (!ctx) return RGPU_ERR_INVALID;
    if (h2d) *h2d = ctx->last_h2d_bytes;
    if (d2h) *d2h = ctx->last_d2h_bytes;
    return RGPU_OK;
}

int rgpu_set_profiling(rgpu_ctx* ctx, int enable) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (enable && !ctx->ev[0])
        for (int i = 0; i < 4; i++) CK(ctx, cudaEventCreate(&ctx->ev[i]));
    ctx->profiling = enable != 0;
    ctx->ev_valid = false;
    return RGPU_OK;
}

int rgpu_last_stage_ms(rgpu_ctx* ctx, float out[3]) {
    if (!ctx || !out) return RGPU_ERR_INVALID;
    if (!ctx->ev_valid) return fail(ctx, RGPU_ERR_INVALID, "no profiled batch (call rgpu_set_profiling(ctx, 1) first)");
    CK(ctx, cudaEventSynchronize(ctx->ev[3]));
    for (int i = 0; i < 3; i++) CK(ctx, cudaEventElapsedTime(&out[i], ctx->ev[i], ctx->ev[i + 1]));
    return RGPU_OK;
}

int rgpu_to_rgba8_dev(rgpu_ctx* ctx, const float* lin_dev, uint8_t* rgba_dev, size_t n_pixels) {
    if (!ctx) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    launch_to_rgba8(reinterpret_cast<const float4*>(lin_dev), reinterpret_cast<uchar4*>(rgba_dev), n_pixels, ctx->stream);
    ctx->n_launches++;
    CK(ctx, cudaGetLastError());
    return RGPU_OK;
}

int rgpu_fill_color_dev(rgpu_ctx* ctx, float* lin_dev, size_t n_pixels, const float color[4]) {
    if (!ctx || !color) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    launch_fill_color(reinterpret_cast<float4*>(lin_dev), n_pixels, make_float4(color[0], color[1], color[2], color[3]), ctx->stream);
    ctx->n_launches++;
    CK(ctx, cudaGetLastError());
    return RGPU_OK;
}

int rgpu_layer_scale_by_mask_dev(rgpu_ctx* ctx, float* lin_dev, size_t lin_origin, size_t lin_stride, const float* mask_dev,
                                 size_t mask_origin, size_t mask_stride, size_t width, size_t height) {
    if (!ctx || ((!lin_dev || !mask_dev) && width * height)) return RGPU_ERR_INVALID;
    if (width > 0xffffffffull || height > 0xffffffffull) return fail(ctx, RGPU_ERR_INVALID, "layer too large");
    CK(ctx, cudaSetDevice(ctx->device));
    launch_scale_by_mask(reinterpret_cast<float4*>(lin_dev) + lin_origin, lin_stride, mask_dev + mask_origin, mask_stride, (uint32_t)width,
                         (uint32_t)height, ctx->stream);
    ctx->n_launches++;
    CK(ctx, cudaGetLastError());
    return RGPU_OK;
}

int rgpu_layer_blend_over_dev(rgpu_ctx* ctx, float* dst_dev, size_t dst_origin, size_t dst_stride, const float* src_dev, size_t src_origin,
                              size_t src_stride, size_t width, size_t height, int use_opacity, float opacity) {
    if (!ctx || ((!dst_dev || !src_dev) && width * height)) return RGPU_ERR_INVALID;
    if (width > 0xffffffffull || height > 0xffffffffull) return fail(ctx, RGPU_ERR_INVALID, "layer too large");
    CK(ctx, cudaSetDevice(ctx->device));
    launch_blend_over(reinterpret_cast<float4*>(dst_dev) + dst_origin, dst_stride, reinterpret_cast<const float4*>(src_dev) + src_origin,
                      src_stride, (uint32_t)width, (uint32_t)height, use_opacity != 0, opacity, ctx->stream);
    ctx->n_launches++;
    CK(ctx, cudaGetLastError());
    return RGPU_OK;
}

int rgpu_download_rgba8(rgpu_ctx* ctx, const float* lin_dev, size_t n_pixels, uint8_t* rgba_host) {
    if (!ctx || ((!lin_dev || !rgba_host) && n_pixels)) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (n_pixels == 0) return RGPU_OK;
    int rc;
    if ((rc = ensure_dev(ctx, ctx->img_f32, 4 * n_pixels))) return rc;  // RGBA8 staging shares the f32 mask scratch
    uchar4* d_rgba = static_cast<uchar4*>(ctx->img_f32.p);
    launch_to_rgba8(reinterpret_cast<const float4*>(lin_dev), d_rgba, n_pixels, ctx->stream);
    ctx->n_launches++;
    CK(ctx, cudaMemcpyAsync(rgba_host, d_rgba, 4 * n_pixels, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RGPU_OK;
}

// ---- trait-level entry points ------------------------------------------------------------------------

int rgpu_flatten(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int close, double* lines_out, size_t cap, size_t* n_out) {
    if (!ctx || !tr || !n_out) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = validate_path(ctx, path);
    if (rc) return rc;
    *n_out = 0;
    if (path->n_segments == 0 || path->n_subpaths == 0) return RGPU_OK;
    rgpu_dpath dp;
    rc = stage_path(ctx, path, &dp);
    if (rc) return rc;
    // flatten only: run count + scan + emit through a throw-away 1x1 mask job so that one code path serves both
    float* d_dummy = nullptr;
    if ((rc = ensure_dev(ctx, ctx->img_f32, 64))) return rc;
    d_dummy = static_cast<float*>(ctx->img_f32.p);
    rgpu_job job;
    std::memset(&job, 0, sizeof(job));
    job.path = &dp;
    std::memcpy(job.tr, tr, sizeof(job.tr));
    job.mode = RGPU_JOB_COVERAGE;
    job.canvas = d_dummy;
    job.row_stride = 1;
    job.width = 1;
    job.height = 1;
    rc = submit_sync(ctx, &job, 1, RGPU_BATCH_INDEPENDENT, close ? 1 : 0, /*ordered_lines=*/true);
    if (rc == RGPU_OK) {
        size_t n = ctx->last_lines;
        *n_out = n;
        if (n > cap || (n && !lines_out)) {
            rc = fail(ctx, RGPU_ERR_CAPACITY, "lines_out too small");
        } else if (n) {
            cudaError_t e = cudaMemcpyAsync(lines_out, ctx->lines.p, n * sizeof(double4), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = RGPU_ERR_CUDA; }
        }
    }
    return rc;
}

static int mask_to_device(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, int mode, size_t width, size_t height,
                          float** d_out) {
    int rc = validate_path(ctx, path);
    if (rc) return rc;
    if (fill_rule != RGPU_NONZERO && fill_rule != RGPU_EVENODD) return fail(ctx, RGPU_ERR_INVALID, "bad fill rule");
    if (width > 0x7ffffff0u || height > 0x7ffffff0u) return fail(ctx, RGPU_ERR_INVALID, "image too large");
    if ((rc = ensure_dev(ctx, ctx->img_f32, sizeof(float) * width * height))) return rc;
    float* d_img = static_cast<float*>(ctx->img_f32.p);
    *d_out = d_img;
    rgpu_dpath dp;
    if (path->n_segments == 0 || path->n_subpaths == 0) {
        CK(ctx, cudaMemsetAsync(d_img, 0, sizeof(float) * width * height, ctx->stream));
        return RGPU_OK;
    }
    static const bool trace = getenv("RGPU_E2E_TRACE") != nullptr;  // timing breakdown on stderr
    const auto t0 = std::chrono::steady_clock::now();
    rc = stage_path(ctx, path, &dp);
    if (rc) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rgpu_job job;
    std::memset(&job, 0, sizeof(job));
    job.path = &dp;
    std::memcpy(job.tr, tr, sizeof(job.tr));
    job.fill_rule = fill_rule;
    job.mode = mode;
    job.canvas = d_img;
    job.row_stride = width;
    job.width = (uint32_t)width;
    job.height = (uint32_t)height;
    rc = submit_sync(ctx, &job, 1, RGPU_BATCH_INDEPENDENT, 1);
    if (trace) {
        const auto t2 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "mask_to_device: stage_path %.3f ms (host item lists + 2 H2D enqueued), submit + kernels + status sync %.3f ms\n", ms(t0, t1),
                     ms(t1, t2));
    }
    return rc;
}

// Run-coded download (compact.cu) of `rows` dense f32 rows of `width` pixels at d_img: device row i becomes row img_row[i] of
// the caller's image (`stride` elements per row, T = float or double).  The class bytes, one literal offset per row and the
// literal segments cross PCIe; the pool's threads rebuild the rows with streaming stores while later literals still copy.
// Returns 1 without touching the image when it declines (small image, RGPU_E2E_RUNCODE=0, or more than half of the segments
// are literals): the caller then copies the dense rows.
extern "C++" {
template <class T>
static int download_runcoded(rgpu_ctx* ctx, const float* d_img, size_t width, size_t rows, const std::vector<size_t>& img_row, T* img, size_t stride) {
    const bool enabled = !(getenv("RGPU_E2E_RUNCODE") && atoi(getenv("RGPU_E2E_RUNCODE")) == 0);  // read per call: A/B inside one process
    if (!enabled || rows * width < ((size_t)4 << 20) || rows > 0x7ffffff0u) return 1;
    // an image of edges was declined a moment ago: the next calls of the same size skip the attempt (0.06 ms each), every 16th looks again
    if (ctx->rc_skip && ctx->rc_skip_w == width && ctx->rc_skip_h == rows) {
        ctx->rc_skip--;
        return 1;
    }
    static const bool trace = getenv("RGPU_E2E_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    int rc;
    const uint32_t segs = runcode_segments(width);
    const size_t n_cls = rows * segs, words = (rows + 1 + 3) & ~(size_t)3;
    if ((rc = ensure_dev(ctx, ctx->rc_buf, sizeof(uint32_t) * 2 * words + n_cls))) return rc;
    if ((rc = ensure_dev(ctx, ctx->scan_temp, scan_temp_bytes((uint32_t)rows + 1)))) return rc;
    if ((rc = ensure_pinned(ctx, ctx->h_rc, ctx->h_rc_cap, sizeof(uint32_t) * words + n_cls))) return rc;
    uint32_t* d_cnt = static_cast<uint32_t*>(ctx->rc_buf.p);
    uint32_t* d_off = d_cnt + words;
    unsigned char* d_cls = reinterpret_cast<unsigned char*>(d_off + words);
    launch_seg_classify(d_img, width, (uint32_t)rows, d_cls, d_cnt, ctx->stream);
    launch_exclusive_scan(d_cnt, d_off, (uint32_t)rows + 1, ctx->scan_temp.p, ctx->scan_temp.cap, ctx->stream);
    ctx->n_launches += 2;
    // (offsets and class bytes are adjacent on both sides: one copy)
    CK(ctx, cudaMemcpyAsync(ctx->h_rc, d_off, sizeof(uint32_t) * words + n_cls, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t* row_off = reinterpret_cast<const uint32_t*>(ctx->h_rc);
    const unsigned char* cls = ctx->h_rc + sizeof(uint32_t) * words;
    const size_t n_lit = row_off[rows];
    const double t_classified = since();
    if (trace && n_lit * 2 > n_cls) std::fprintf(stderr, "download_runcoded: %zu of %zu segments literal: declined after %.3f ms\n", n_lit, n_cls, t_classified);
    if (n_lit * 2 > n_cls) {  // an image of edges: dense copies are the shorter way
        ctx->rc_skip = 15;
        ctx->rc_skip_w = width;
        ctx->rc_skip_h = rows;
        return 1;
    }
    if ((rc = ensure_dev(ctx, ctx->rc_lits, std::max<size_t>(sizeof(float) * 64 * n_lit, 16)))) return rc;
    if ((rc = ensure_pinned(ctx, ctx->h_lits, ctx->h_lits_cap, std::max<size_t>(64 * n_lit, 16)))) return rc;
    float* d_lits = static_cast<float*>(ctx->rc_lits.p);
    if (n_lit) {
        launch_seg_emit(d_img, width, (uint32_t)rows, d_cls, d_off, d_lits, ctx->stream);
        ctx->n_launches += 1;
    }
    ensure_pool(ctx);
    // the literals come down in pieces of rows, so that the rows of a piece are rebuilt while the next piece is copied
    const size_t pieces = std::min<size_t>(32, std::max<size_t>(1, rows / 64));
    while (ctx->chunk_ev.size() < pieces) {
        cudaEvent_t e;
        CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_ev.push_back(e);
    }
    for (size_t p = 0; p < pieces; p++) {
        const size_t ra = rows * p / pieces, rb = rows * (p + 1) / pieces;
        const size_t la = row_off[ra], lb = row_off[rb];
        if (lb > la) CK(ctx, cudaMemcpyAsync(ctx->h_lits + 64 * la, d_lits + 64 * la, sizeof(float) * 64 * (lb - la), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaEventRecord(ctx->chunk_ev[p], ctx->stream));
    }
    const unsigned workers = ctx->pool->size();
    const float* lits = ctx->h_lits;
    const size_t* rowmap = img_row.data();
    struct PoolDrain {  // no task may outlive the row map, whichever way the call ends
        rgpu::HostPool* p;
        ~PoolDrain() { p->wait(); }
    } pool_drain{ctx->pool.get()};
    for (size_t p = 0; p < pieces; p++) {
        const size_t ra = rows * p / pieces, rb = rows * (p + 1) / pieces;
        CK(ctx, cudaEventSynchronize(ctx->chunk_ev[p]));
        const size_t parts = std::min<size_t>(workers, rb - ra);
        for (size_t q = 0; q < parts; q++) {
            const size_t a = ra + (rb - ra) * q / parts, b = ra + (rb - ra) * (q + 1) / parts;
            ctx->pool->submit([=] {
                for (size_t i = a; i < b; i++) {
                    T* drow = img + rowmap[i] * stride;
                    if (sizeof(T) == 4) rgpu::expand_runs_f32(cls + i * segs, segs, width, lits + 64 * (size_t)row_off[i], reinterpret_cast<float*>(drow));
                    else rgpu::expand_runs_f64(cls + i * segs, segs, width, lits + 64 * (size_t)row_off[i], reinterpret_cast<double*>(drow));
                }
                rgpu::host_store_fence();
            });
        }
    }
    const double t_copied = since();
    ctx->pool->wait();
    ctx->last_d2h_bytes = sizeof(uint32_t) * words + n_cls + sizeof(float) * 64 * n_lit;
    if (trace)
        std::fprintf(stderr, "download_runcoded: %zu x %zu, %zu of %zu segments literal (%.1f MB over PCIe instead of %.1f MB); classified %.3f ms, "
                     "literals copied %.3f ms, rows rebuilt %.3f ms (%s)\n", width, rows, n_lit, n_cls, ctx->last_d2h_bytes / 1e6,
                     rows * width * 4 / 1e6, t_classified, t_copied, since(), rgpu::host_simd_name());
    return RGPU_OK;
}
}  // extern "C++"

int rgpu_mask_f32(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, float* img, size_t width, size_t height) {
    if (!ctx || !tr || (!img && width * height)) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (width == 0 || height == 0) return RGPU_OK;
    float* d = nullptr;
    int rc = mask_to_device(ctx, path, tr, fill_rule, RGPU_JOB_MASK, width, height, &d);
    if (rc) return rc;
    {   // large masks come down run-coded (constant segments as class bytes, rebuilt by host threads)
        std::vector<size_t> img_row(height);
        for (size_t y = 0; y < height; y++) img_row[y] = y;
        const uint64_t h2d = ctx->last_h2d_bytes;
        rc = download_runcoded<float>(ctx, d, width, height, img_row, img, width);
        ctx->last_h2d_bytes = h2d;
        if (rc <= 0) return rc;
    }
    CK(ctx, cudaMemcpyAsync(img, d, sizeof(float) * width * height, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->last_d2h_bytes = sizeof(float) * width * height;
    return RGPU_OK;
}

int rgpu_coverage_f32(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, float* out, size_t width, size_t height) {
    if (!ctx || !tr || (!out && width * height)) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (width == 0 || height == 0) return RGPU_OK;  // src/rasterize.rs:320-322
    float* d = nullptr;
    int rc = mask_to_device(ctx, path, tr, fill_rule, RGPU_JOB_COVERAGE, width, height, &d);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(out, d, sizeof(float) * width * height, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RGPU_OK;
}

// f32 -> f64 row with non-temporal stores: a plain store stream would first read every destination line for
// ownership (134 MB of extra DRAM reads on a 4096^2 mask), which is what bounds the host side of rgpu_mask.
static inline void widen_row(const float* __restrict__ src, double* __restrict__ dst, size_t n) { rgpu::widen_row_simd(src, dst, n); }

// f32 device image -> strided f64 host image.  Two producers fill the caller's image at once:
//   * the BOTTOM rows cross PCIe as f32 in row chunks; as soon as a chunk has landed in pinned staging the pool widens
//     it into the caller's image while the next chunks are still in flight (the caller's memory may be pageable: only
//     host threads write it);
//   * when the caller's image is pinned and dense along x, the TOP rows are widened on the device and DMA'd straight
//     into place as f64 behind the f32 chunks, so the copy engine keeps working while the host threads catch up.
// The split adapts from call to call to whichever side finished last (host widening bandwidth differs a lot between
// hosts: 134 MB of f64 stores per 4096^2 mask).  Measured and rejected: a lossless row encoding of the mask (bit masks of
// exact 0 / 1 per 32-pixel word + packed literals) expanded by the host threads — a quarter of the pixels of a 4096^2 mask
// are literals (covered pixels are 1 - k * 2^-24 as often as exactly 1), the encode kernel costs 0.3 ms, and the call got slower.
static int download_widen(rgpu_ctx* ctx, const float* d_img, size_t w, size_t h, double* dst, rgpu_shape shape) {
    int rc;
    if ((rc = ensure_stage(ctx, sizeof(float) * w * h))) return rc;
    ensure_pool(ctx);
    // rows [0, h_dev) go the device-widened way
    size_t h_dev = 0;
    if (shape.col_stride == 1 && h >= 64) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
            h_dev = std::min(h - 1, (size_t)((double)h * ctx->widen_dev_frac));
        } else {
            cudaGetLastError();  // pageable memory reports an error on older runtimes
        }
    }
    if (h_dev) {
        if ((rc = ensure_dev(ctx, ctx->img_f64, sizeof(double) * w * h_dev))) return rc;
        launch_f32_to_f64(d_img, static_cast<double*>(ctx->img_f64.p), w * h_dev, ctx->stream);
        ctx->n_launches++;
    }
    float* stage = static_cast<float*>(ctx->h_stage);
    const size_t target_rows = std::max<size_t>(1, (size_t)(4u << 20) / (w * sizeof(float)));  // ~4 MB per chunk
    const size_t n_chunks = (h - h_dev + target_rows - 1) / target_rows;
    while (ctx->chunk_ev.size() < n_chunks + 1) {
        cudaEvent_t e;
        CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_ev.push_back(e);
    }
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t r0 = h_dev + c * target_rows, r1 = std::min(h, r0 + target_rows);
        CK(ctx, cudaMemcpyAsync(stage + r0 * w, d_img + r0 * w, sizeof(float) * w * (r1 - r0), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaEventRecord(ctx->chunk_ev[c], ctx->stream));
    }
    if (h_dev) {
        CK(ctx, cudaMemcpy2DAsync(dst, shape.row_stride * sizeof(double), ctx->img_f64.p, w * sizeof(double), w * sizeof(double), h_dev,
                                  cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaEventRecord(ctx->chunk_ev[n_chunks], ctx->stream));
    }
    ctx->last_d2h_bytes = sizeof(float) * w * (h - h_dev) + sizeof(double) * w * h_dev;
    const unsigned workers = ctx->pool->size();
    const size_t rs = shape.row_stride, cs = shape.col_stride;
    static const bool trace = getenv("RGPU_E2E_TRACE") != nullptr;  // timing breakdown of this function on stderr
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    double t_first = 0, t_last = 0;
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t r0 = h_dev + c * target_rows, r1 = std::min(h, r0 + target_rows);
        CK(ctx, cudaEventSynchronize(ctx->chunk_ev[c]));
        if (c == 0) t_first = since();
        t_last = since();
        const size_t rows = r1 - r0;
        const size_t parts = std::min<size_t>(workers, rows);
        for (size_t p = 0; p < parts; p++) {
            const size_t a = r0 + rows * p / parts, b = r0 + rows * (p + 1) / parts;
            ctx->pool->submit([=] {
                for (size_t y = a; y < b; y++) {
                    const float* srow = stage + y * w;
                    double* drow = dst + y * rs;
                    if (cs == 1) {
                        widen_row(srow, drow, w);
                    } else {
                        for (size_t x = 0; x < w; x++) drow[x * cs] = (double)srow[x];
                    }
                }
#if defined(__SSE2__)
                _mm_sfence();
#endif
            });
        }
    }
    if (h_dev) {
        // which producer finishes last?  (the f64 DMA is queued behind the f32 chunks)  Move the split towards the other.
        ctx->pool->wait();
        const double t_pool = since();
        const bool host_was_last = cudaEventQuery(ctx->chunk_ev[n_chunks]) == cudaSuccess;
        CK(ctx, cudaEventSynchronize(ctx->chunk_ev[n_chunks]));
        if (trace)
            fprintf(stderr, "download_widen: dev share %.2f, first f32 chunk %.3f ms, last %.3f ms, pool done %.3f ms, f64 dma done %.3f ms\n",
                    ctx->widen_dev_frac, t_first, t_last, t_pool, since());
        if (host_was_last) ctx->widen_dev_frac = std::min(0.6, ctx->widen_dev_frac + 0.02);
        else ctx->widen_dev_frac = std::max(0.0, ctx->widen_dev_frac - 0.02);
    } else {
        ctx->pool->wait();
    }
    return RGPU_OK;
}

int rgpu_mask(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, double* img, rgpu_shape shape) {
    if (!ctx || !tr) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    size_t w = shape.width, h = shape.height;
    if (w == 0 || h == 0) return RGPU_OK;
    if (!img) return RGPU_ERR_INVALID;
    float* d = nullptr;
    int rc = mask_to_device(ctx, path, tr, fill_rule, RGPU_JOB_MASK, w, h, &d);
    if (rc) return rc;
    if (shape.col_stride == 1) {  // large masks come down run-coded and are rebuilt as f64 by host threads (download_runcoded)
        std::vector<size_t> img_row(h);
        for (size_t y = 0; y < h; y++) img_row[y] = y;
        const uint64_t h2d = ctx->last_h2d_bytes;
        rc = download_runcoded<double>(ctx, d, w, h, img_row, img + shape.start, shape.row_stride);
        ctx->last_h2d_bytes = h2d;
        if (rc <= 0) return rc;
    }
    return download_widen(ctx, d, w, h, img + shape.start, shape);
}

int rgpu_mask_iter(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], size_t width, size_t height, int fill_rule,
                   rgpu_pixel* out, size_t cap, size_t* n_out) {
    if (!ctx || !tr || !n_out) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    *n_out = 0;
    if (width == 0 || height == 0) return RGPU_OK;
    float* d = nullptr;
    int rc = mask_to_device(ctx, path, tr, fill_rule, RGPU_JOB_COVERAGE, width, height, &d);
    if (rc) return rc;
    // the yielded pixels are compacted on the device (compact.cu): count per block of pixels, scan, emit in row-major order —
    // only the records cross PCIe (the dense canvas used to, followed by a scan on one host thread)
    static_assert(sizeof(rgpu_pixel) == 24, "rgpu_pixel layout (compact.cu PixelRec)");
    const size_t npx = width * height;
    if (npx > 0xffffffffull) return fail(ctx, RGPU_ERR_INVALID, "canvas too large for a pixel list (more than 2^32 pixels)");
    const uint32_t nb = pixel_blocks(npx);
    const size_t offs_at = ((size_t)nb + 1 + 3) & ~(size_t)3;  // the scan moves whole uint4s: both arrays 16-byte aligned
    if ((rc = ensure_dev(ctx, ctx->px_counts, sizeof(uint32_t) * (offs_at + (size_t)nb + 1 + 4)))) return rc;
    if ((rc = ensure_dev(ctx, ctx->scan_temp, scan_temp_bytes(nb + 1)))) return rc;
    if ((rc = ensure_stage(ctx, 16))) return rc;
    uint32_t* d_counts = static_cast<uint32_t*>(ctx->px_counts.p);
    uint32_t* d_offs = d_counts + offs_at;
    launch_pixel_count(d, npx, d_counts, ctx->stream);
    launch_exclusive_scan(d_counts, d_offs, nb + 1, ctx->scan_temp.p, ctx->scan_temp.cap, ctx->stream);
    CK(ctx, cudaMemcpyAsync(ctx->h_stage, d_offs + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_launches += 2;
    const size_t n = *static_cast<const uint32_t*>(ctx->h_stage);
    const size_t take = out ? std::min(n, cap) : 0;
    ctx->last_d2h_bytes = sizeof(uint32_t);
    if (take) {
        if ((rc = ensure_dev(ctx, ctx->px_out, sizeof(rgpu_pixel) * take))) return rc;
        launch_pixel_emit(d, npx, width, d_offs, ctx->px_out.p, take, ctx->stream);
        ctx->n_launches += 1;
        CK(ctx, cudaMemcpyAsync(out, ctx->px_out.p, sizeof(rgpu_pixel) * take, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->last_d2h_bytes += sizeof(rgpu_pixel) * take;
    }
    *n_out = n;
    if (n > cap) return fail(ctx, RGPU_ERR_CAPACITY, "pixel buffer too small");
    return RGPU_OK;
}

int rgpu_fill(rgpu_ctx* ctx, const rgpu_path* path, const double tr[6], int fill_rule, const rgpu_paint* paint,
              const double* path_bbox, float* img, rgpu_shape shape) {
    if (!ctx || !tr || !paint) return RGPU_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    size_t w = shape.width, h = shape.height;
    if (w == 0 || h == 0) return RGPU_OK;
    if (!img) return RGPU_ERR_INVALID;
    int rc = validate_path(ctx, path);
    if (rc) return rc;
    if (path->n_segments == 0 || path->n_subpaths == 0) return RGPU_OK;
    if (paint->n_stops > RGPU_MAX_STOPS) return fail(ctx, RGPU_ERR_INVALID, "too many gradient stops");
    if ((rc = ensure_dev(ctx, ctx->img_lin, sizeof(float4) * w * h))) return rc;
    float4* d_img = static_cast<float4*>(ctx->img_lin.p);
    float* dst = img + 4 * shape.start;
    bool dense_rows = shape.col_stride == 1;
    // Only pixels inside the bounding box of the transformed control points (the curve lies in their convex hull) can be
    // covered, so only that rectangle of the image has to cross PCIe, both ways; the kernels still see the whole view
    // (clipping at the view's edges is part of the reference's arithmetic).  Parts left of the view fold onto column 0
    // and parts beyond the right edge onto the last column: the clamps below keep those columns in.
    size_t rx0 = 0, rx1 = w, ry0 = 0, ry1 = h;
    {
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        bool finite = true;
        for (uint32_t i = 0; i < path->n_points; i++) {
            const double px = path->points[2 * i], py = path->points[2 * i + 1];
            const double x = px * tr[0] + py * tr[1] + tr[2], y = px * tr[3] + py * tr[4] + tr[5];
            finite = finite && std::isfinite(x) && std::isfinite(y);
            xmin = std::min(xmin, x); xmax = std::max(xmax, x);
            ymin = std::min(ymin, y); ymax = std::max(ymax, y);
        }
        if (finite && path->n_points) {
            auto clampi = [](double v, size_t hi) { return (size_t)std::min<double>(std::max(v, 0.0), (double)hi); };
            rx0 = clampi(std::floor(xmin) - 2.0, w); rx1 = clampi(std::ceil(xmax) + 3.0, w);
            ry0 = clampi(std::floor(ymin) - 1.0, h); ry1 = clampi(std::ceil(ymax) + 2.0, h);
            if (rx1 <= rx0 || ry1 <= ry0) rx0 = rx1 = ry0 = ry1 = 0;  // nothing of the view can change
        }
    }
    const size_t rw = rx1 - rx0, rh = ry1 - ry0;
    // host image -> device canvas
    if (dense_rows) {
        if (rw && rh)
            CK(ctx, cudaMemcpy2DAsync(d_img + ry0 * w + rx0, w * sizeof(float4), dst + 4 * (ry0 * shape.row_stride + rx0),
                                      shape.row_stride * sizeof(float4), rw * sizeof(float4), rh, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // a view with a column stride: the same rectangle, gathered into (and later scattered from) a packed pinned staging
        if ((rc = ensure_stage(ctx, std::max<size_t>(sizeof(float4) * rw * rh, 16)))) return rc;
        float4* st = static_cast<float4*>(ctx->h_stage);
        for (size_t y = 0; y < rh; y++)
            for (size_t x = 0; x < rw; x++) std::memcpy(&st[y * rw + x], dst + 4 * ((ry0 + y) * shape.row_stride + (rx0 + x) * shape.col_stride), 16);
        if (rw && rh)
            CK(ctx, cudaMemcpy2DAsync(d_img + ry0 * w + rx0, w * sizeof(float4), st, rw * sizeof(float4), rw * sizeof(float4), rh, cudaMemcpyHostToDevice,
                                      ctx->stream));
    }
    rgpu_dpath dp;
    rc = stage_path(ctx, path, &dp);
    if (rc) return rc;
    rgpu_job job;
    std::memset(&job, 0, sizeof(job));
    job.path = &dp;
    std::memcpy(job.tr, tr, sizeof(job.tr));
    job.fill_rule = fill_rule;
    job.mode = RGPU_JOB_FILL;
    job.paint = paint;
    job.path_bbox = path_bbox;
    job.canvas = d_img;
    job.row_stride = w;
    job.width = (uint32_t)w;
    job.height = (uint32_t)h;
    rc = submit_sync(ctx, &job, 1, RGPU_BATCH_ORDERED, 1);
    if (rc == RGPU_ERR_WINDING && ctx->fix_shift > kFixShiftWide) {
        // the fill has blended into the device copy with wrapped windings: restore it from the caller's image and repeat in Q13.18
        if (dense_rows) {
            if (rw && rh)
                CK(ctx, cudaMemcpy2DAsync(d_img + ry0 * w + rx0, w * sizeof(float4), dst + 4 * (ry0 * shape.row_stride + rx0),
                                          shape.row_stride * sizeof(float4), rw * sizeof(float4), rh, cudaMemcpyHostToDevice, ctx->stream));
        } else if (rw && rh) {
            CK(ctx, cudaMemcpy2DAsync(d_img + ry0 * w + rx0, w * sizeof(float4), ctx->h_stage, rw * sizeof(float4), rw * sizeof(float4), rh,
                                      cudaMemcpyHostToDevice, ctx->stream));
        }
        const int keep = ctx->fix_shift;
        ctx->fix_shift = kFixShiftWide;
        rc = submit_sync(ctx, &job, 1, RGPU_BATCH_ORDERED, 1);
        ctx->fix_shift = keep;
    }
    if (rc) return rc;
    if (dense_rows) {
        if (rw && rh)
            CK(ctx, cudaMemcpy2DAsync(dst + 4 * (ry0 * shape.row_stride + rx0), shape.row_stride * sizeof(float4), d_img + ry0 * w + rx0,
                                      w * sizeof(float4), rw * sizeof(float4), rh, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        float4* st = static_cast<float4*>(ctx->h_stage);
        if (rw && rh)
            CK(ctx, cudaMemcpy2DAsync(st, rw * sizeof(float4), d_img + ry0 * w + rx0, w * sizeof(float4), rw * sizeof(float4), rh, cudaMemcpyDeviceToHost,
                                      ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t y = 0; y < rh; y++)
            for (size_t x = 0; x < rw; x++) std::memcpy(dst + 4 * ((ry0 + y) * shape.row_stride + (rx0 + x) * shape.col_stride), &st[y * rw + x], 16);
    }
    return RGPU_OK;
}

}  // extern "C"

#include "multi.inl"
#include "stroke_host.inl"
#include "parse_host.inl"
