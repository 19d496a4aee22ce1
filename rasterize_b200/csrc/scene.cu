// Scene compositor — every Fill of a layer in ONE launch, one CTA per LAYER tile.
//
// Replaces the per-node loop of `Pipeline::render_rec`'s Fill arm (reference src/scene.rs:397-435: for every Fill node
// `path.fill(rasterizer, align * tr, fill_rule, paint, layer.view_mut(..))`) together with `Layer::new`'s background
// (src/scene.rs:483-501) and the RGBA8 export that follows `Scene::render` in every CLI run (src/color.rs:164-175).
//
// The ordered batch of raster.cu issues one launch per fill because fills blend in order onto one canvas.  Here the order
// is kept INSIDE a CTA instead: a CTA owns a 256 x 8 tile of the layer, keeps its LinColor pixels in shared memory
// (32 KB), and walks the fills in submission order — accumulate the fill's lines of this tile (same fixed-point
// signed-difference cells as raster.cu), carry-in from the tile to the left, row scan, fill rule, `Paint::at`,
// `with_alpha`, `blend_over` into the shared tile.  The layer is read at most once and written once, however many fills
// overlap (SURVEY §8d "(s)" bytes: 16 B x W x H, + 4 B with the RGBA8 export), and tiles the fills do not touch cost
// nothing but the store.
//
// For this the jobs' tile grids are aligned with the layer's (JobDev::ox/oy/sc0/sb0): the flatten kernel bins a line
// into the job tiles (job, band, chunk) exactly as for raster.cu, and job tile (b, c) IS layer tile (sb0 + b, sc0 + c).
// The carry between the chunks of a band is a chained inclusive prefix per (job, tile, row) in `tile_state`: layer tiles
// are taken in chunk-major ticket order, so the tile a CTA waits on was started a whole column of bands earlier.
#include "raster_device.cuh"

#include <cstdlib>

namespace rgpu {

namespace {

using namespace rs;

constexpr int kScH = 8;
constexpr int kScSpanCap = 208;       // per-warp span list (lane << 3 | row)
constexpr int kScPre = 32;            // fills of a tile whose bin counter / carry words are fetched up front
using ScSpanT = unsigned char;
constexpr size_t kScPaintBytes = 2048;  // the paint sits at the start of the piece constants ...
static_assert(kScPaintBytes >= sizeof(PaintDev), "the paint is staged over the piece constants");
template <int CW, int THREADS>
struct ScCfg {
    static constexpr int kWarps = THREADS / 32;
    static constexpr int kL = CW / 32;              // columns per lane in the row scan
    static constexpr int kPix = CW * kScH;          // pixels of a tile
    static constexpr int kPasses = kPix / THREADS;  // pixels per thread
    static constexpr size_t kColorBytes = sizeof(float4) * kPix;
    static constexpr size_t kCellBytes = sizeof(int) * kPix;
    // piece constants of the accumulation; reused for the paint + the covered-pixel list while compositing
    static constexpr size_t kPieceBytes = (sizeof(double) * 4 * THREADS > kScPaintBytes + sizeof(unsigned short) * kPix)
                                              ? sizeof(double) * 4 * THREADS : kScPaintBytes + sizeof(unsigned short) * kPix;
    static constexpr size_t kSmem = kColorBytes + kCellBytes + kPieceBytes + sizeof(ScSpanT) * kScSpanCap * kWarps;
    static constexpr int kMinBlocks = (THREADS >= 512) ? 2 : 4;  // 64 registers per thread
    static_assert(CW % 128 == 0 && kPix % THREADS == 0 && kPix <= 65536 && kWarps >= kScH, "tile shape");
};

template <bool EVENODD, int kScL>
__device__ __forceinline__ void scan_row_inplace(int* bc, int acc, int lane, const Fix fix, const bool wcheck, Status* __restrict__ status) {
    int v[kScL];
#pragma unroll
    for (int i = 0; i < kScL / 4; i++) {
        const int4 q = *reinterpret_cast<const int4*>(bc + swz<true>(lane * kScL + i * 4));
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int i = 1; i < kScL; i++) v[i] += v[i - 1];
    int incl = v[kScL - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nb;
    }
    const int base = acc + incl - v[kScL - 1];
    if (!EVENODD && wcheck) {  // winding guard, see raster.cu scan_rows
        bool risk = false;
#pragma unroll
        for (int i = 0; i < kScL / 4; i++) risk = risk || winding_risk(base + v[4 * i], base + v[4 * i + 1], base + v[4 * i + 2], base + v[4 * i + 3], fix);
        if (risk) status->winding_flag = 1u;
    }
#pragma unroll
    for (int i = 0; i < kScL / 4; i++) {
        const float4 cv = make_float4(coverage_from_fixed<EVENODD>(base + v[4 * i], fix), coverage_from_fixed<EVENODD>(base + v[4 * i + 1], fix),
                                      coverage_from_fixed<EVENODD>(base + v[4 * i + 2], fix), coverage_from_fixed<EVENODD>(base + v[4 * i + 3], fix));
        *reinterpret_cast<float4*>(bc + swz<true>(lane * kScL + i * 4)) = cv;
    }
}

template <int kScW, int kScThreads>
__global__ void __launch_bounds__(kScThreads, (ScCfg<kScW, kScThreads>::kMinBlocks))
scene_kernel(const JobDev* __restrict__ jobs, uint32_t n_jobs, const PaintDev* __restrict__ paints, uint32_t* __restrict__ tile_offs,
             uint32_t bin_cap, const double4* __restrict__ bin_lines, unsigned long long* __restrict__ tile_state, uint32_t epoch,
             uint32_t* __restrict__ ticket, Status* status, const SceneArgs sc) {
    using Cfg = ScCfg<kScW, kScThreads>;
    constexpr int kScL = Cfg::kL, kScWarps = Cfg::kWarps;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* color = reinterpret_cast<float4*>(smem_raw);
    int* cells = reinterpret_cast<int*>(smem_raw + Cfg::kColorBytes);
    double* p_ax = reinterpret_cast<double*>(smem_raw + Cfg::kColorBytes + Cfg::kCellBytes);
    double* p_ay = p_ax + kScThreads;
    double* p_by = p_ay + kScThreads;
    double* p_dxdy = p_by + kScThreads;
    ScSpanT* spans_all = reinterpret_cast<ScSpanT*>(smem_raw + Cfg::kColorBytes + Cfg::kCellBytes + Cfg::kPieceBytes);
    const PaintDev& s_paint = *reinterpret_cast<const PaintDev*>(p_ax);
    unsigned short* cov_list = reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned char*>(p_ax) + kScPaintBytes);
    __shared__ int carry[kScH], rowtot[kScH], row_touched[kScH], row_live[kScH];
    __shared__ float row_const[kScH];
    __shared__ uint32_t s_ticket, s_count, s_bad, s_ncov;
    __shared__ uint32_t pre_cnt[kScPre];
    __shared__ unsigned long long pre_state[kScPre][kScH];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t n_tiles = sc.n_bands * sc.n_chunks;
    const uint32_t t = s_ticket;
    // chunk-major order: the left neighbour of a tile started a column of bands earlier
    const int C = (int)(t / sc.n_bands), B = (int)sc.band_order[t - (uint32_t)C * sc.n_bands];
    const int X0 = C * kScW, Y0 = B * kScH;
    const int tw = min(kScW, (int)sc.width - X0), th = min(kScH, (int)sc.height - Y0);
    // The tile's pixels: `Layer::new` background, or what the layer holds (fills queued after a clip / opacity group).
    // Everything launched before the flatten kernel is complete, so this overlaps the flatten kernel's tail.
    {
        const float4 bg = make_float4(sc.bg[0], sc.bg[1], sc.bg[2], sc.bg[3]);
#pragma unroll
        for (int k = 0; k < Cfg::kPasses; k++) {
            const int p = k * kScThreads + tid, r = p / kScW, col = p % kScW;
            float4 c = bg;
            if (!sc.fresh) {
                c = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < th && col < tw) c = sc.layer[(size_t)(Y0 + r) * sc.width + X0 + col];
            }
            color[p] = c;
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
        const volatile uint32_t* flags = reinterpret_cast<const volatile uint32_t*>(status);
        s_bad = flags[0] | flags[1] | flags[2] | flags[3];  // nan, depth, lines_overflow, refs_overflow
        if (t == n_tiles - 1) *ticket = 0u;  // every ticket of this launch is drawn: leave the counter clean
    }
    __syncthreads();
    if (s_bad) {
        // the host re-runs (or reports): leave the layer as it was and the tile counters clean
        for (uint32_t k = sc.band_offs[B] + tid; k < sc.band_offs[B + 1]; k += kScThreads) {
            const JobDev& job = jobs[sc.band_jobs[k]];
            const int b = B - job.sb0, c = C - job.sc0;
            if (b >= 0 && b < (int)job.n_bands && c >= 0 && c < (int)job.n_chunks) tile_offs[job.tile_begin + (uint32_t)b * job.n_chunks + (uint32_t)c] = 0u;
        }
        return;
    }
    const unsigned long long ep = (unsigned long long)epoch << 34;
    bool dirty = sc.fresh != 0;
    ScSpanT* spans = spans_all + warp * kScSpanCap;

    const uint32_t k_begin = sc.band_offs[B], k_end = sc.band_offs[B + 1];
    // One memory round trip for the whole walk instead of two per fill: the bin counters of this tile's first kScPre fills
    // and the state words of their left neighbours are fetched together up front (in chunk-major order the neighbours
    // have normally published long ago; a word that is not ready yet is polled when its fill comes up).
    const uint32_t n_pre = min(k_end - k_begin, (uint32_t)kScPre);
    for (uint32_t i = tid; i < n_pre * (kScH + 1); i += kScThreads) {
        const uint32_t kk = i / (kScH + 1), w = i - kk * (kScH + 1);
        const JobDev& job = jobs[sc.band_jobs[k_begin + kk]];
        const int b = B - job.sb0, c = C - job.sc0;
        const bool mine = b >= 0 && b < (int)job.n_bands && c >= 0 && c < (int)job.n_chunks;
        const uint32_t tile = job.tile_begin + (uint32_t)b * job.n_chunks + (uint32_t)c;
        if (w == 0) {
            uint32_t n = 0;
            if (mine) {
                n = min(tile_offs[tile], bin_cap);
                tile_offs[tile] = 0u;  // self-cleaning counters, as in raster.cu
            }
            pre_cnt[kk] = n;
        } else {
            pre_state[kk][w - 1] = (mine && c > 0) ? ld_state(tile_state + (size_t)(tile - 1u) * kStateRows + (w - 1)) : 0ull;
        }
    }
    __syncthreads();
    for (uint32_t k = k_begin; k < k_end; k++) {
        const JobDev& job = jobs[sc.band_jobs[k]];
        const int b = B - job.sb0, c = C - job.sc0;
        if (b < 0 || b >= (int)job.n_bands || c < 0 || c >= (int)job.n_chunks) continue;  // uniform for the CTA
        const uint32_t tile = job.tile_begin + (uint32_t)b * job.n_chunks + (uint32_t)c;
        const bool chained = job.n_chunks > 1;
        // the descriptor fields the loops below use, read once (the compiler would reload them around barriers and atomics)
        const int job_wout = job.width_out, job_paint = job.paint_index;
        const bool job_evenodd = job.rule == 1;
        const uint32_t kk = k - k_begin;
        if (tid == 0) {
            if (kk < n_pre) {
                s_count = pre_cnt[kk];
            } else {
                s_count = min(tile_offs[tile], bin_cap);
                tile_offs[tile] = 0u;
            }
        }
        if (tid == 32) s_ncov = 0u;
        if (tid < kScH) {
            int cin = 0;
            if (c > 0) {  // inclusive prefix of the job's tile to the left (published by the CTA of layer tile (B, C - 1))
                const unsigned long long* ps = tile_state + (size_t)(tile - 1u) * kStateRows + tid;
                unsigned long long v = (kk < n_pre) ? pre_state[kk][tid] : ld_state(ps);
                while ((uint32_t)(v >> 34) != epoch || (v & (3ull << 32)) != kFlagPrefix) v = ld_state(ps);
                cin = (int)(uint32_t)v;
            }
            carry[tid] = cin;
            rowtot[tid] = 0;
            row_touched[tid] = 0;
        }
        __syncthreads();
        const uint32_t cnt = s_count;
        if (cnt == 0) {
            int any = 0;
#pragma unroll
            for (int r = 0; r < kScH; r++) any |= carry[r];
            if (tid < kScH && chained) st_state(tile_state + (size_t)tile * kStateRows + tid, ep | kFlagPrefix | (unsigned long long)(uint32_t)carry[tid]);
            if (!any) {  // the fill's window covers this tile, its geometry does not
                __syncthreads();
                continue;
            }
        }
        TileGeom g;
        g.row0 = b * kScH - job.oy;  // job rows of this tile: may start above the window (rows < 0 hold nothing)
        g.row1 = min(g.row0 + kScH, job.height);
        g.cx0 = c * kScW - job.ox;   // likewise for columns
        g.wc = job.clamp_w;
        g.wci = (int)g.wc;
        g.tile_end = min(g.cx0 + kScW, g.wci + 1);
        g.pitch = kScW;
        const Fix fix = make_fix(job.fix_shift);
        g.fix_scale = fix.scale;
        if (cnt) {
            {
                const int4 z = make_int4(0, 0, 0, 0);
                int4* c4 = reinterpret_cast<int4*>(cells);
#pragma unroll
                for (int i = tid; i < kScH * kScW / 4; i += kScThreads) c4[i] = z;
            }
            __syncthreads();
            const uint32_t rbeg = tile * bin_cap, rend = rbeg + cnt;
            const uint32_t per = (cnt + kScWarps - 1) / kScWarps;
            const uint32_t wbeg = rbeg + (uint32_t)warp * per, wend = min(wbeg + per, rend);
            for (uint32_t r0 = wbeg; r0 < wend; r0 += 32) {
                const uint32_t r = r0 + lane;
                const bool valid = r < wend;
                const double4 l = valid ? bin_lines[r] : make_double4(0, 0, 0, 0);
                warp_accumulate_round<true, 3, kScSpanCap, ScSpanT>(l, valid, g, cells, rowtot, row_touched, p_ax, p_ay, p_by, p_dxdy, spans, tid);
            }
            __syncthreads();
            if (tid < kScH && chained)
                st_state(tile_state + (size_t)tile * kStateRows + tid, ep | kFlagPrefix | (unsigned long long)(uint32_t)(carry[tid] + rowtot[tid]));
        }
        // the piece constants are dead: stage the paint over them
        if (job_paint >= 0) {
            const int* src = reinterpret_cast<const int*>(&paints[job_paint]);
            int* dst = reinterpret_cast<int*>(p_ax);
            for (int i = tid; i < (int)(sizeof(PaintDev) / 4); i += kScThreads) dst[i] = src[i];
        }
        // row scan + fill rule: one warp per row; coverage stays in the cells (as floats) for the compositing pass
        if (warp < kScH) {
            const int r = warp;
            const int acc = carry[r];
            const bool wcheck = !job_evenodd && (abs(acc) >> job.fix_shift) + (int)cnt >= (fix.guard >> job.fix_shift);
            if (wcheck && lane == 0 && abs(acc) >= fix.guard) status->winding_flag = 1u;
            if (row_touched[r]) {
                if (job_evenodd) scan_row_inplace<true, kScL>(cells + r * kScW, acc, lane, fix, false, status);
                else scan_row_inplace<false, kScL>(cells + r * kScW, acc, lane, fix, wcheck, status);
                if (lane == 0) row_live[r] = 1;
            } else if (lane == 0) {  // no line touched this row of the tile: constant coverage
                const float cv = job_evenodd ? coverage_from_fixed<true>(acc, fix) : coverage_from_fixed<false>(acc, fix);
                row_const[r] = cv;
                row_live[r] = cv >= 1e-6f;
            }
        }
        __syncthreads();
        // paint + composite (src/rasterize.rs:103-115).  `Paint::at` costs hundreds of instructions per pixel, so the tile's
        // covered pixels (alpha >= 1e-6: mask_iter drops the rest, src/rasterize.rs:348) are first compacted into a list —
        // row segments stay contiguous — and then dealt out evenly: every warp works with all lanes, whatever the shape.
        const float* covs = reinterpret_cast<const float*>(cells);
        {
            unsigned mine = 0;  // bit k: this thread's pixel of pass k is covered
            int cnt_before = 0, my_pos[Cfg::kPasses];
            const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
            for (int k = 0; k < Cfg::kPasses; k++) {
                const int p = k * kScThreads + tid, r = p / kScW, col = p % kScW;
                const int x = g.cx0 + col, y = g.row0 + r;
                bool cov = x >= 0 && x < job_wout && y >= 0 && y < g.row1 && row_live[r];
                if (cov) cov = (row_touched[r] ? covs[r * kScW + swz<true>(col)] : row_const[r]) >= 1e-6f;
                const unsigned bal = __ballot_sync(0xffffffffu, cov);
                my_pos[k] = cnt_before + __popc(bal & lt_mask);
                cnt_before += __popc(bal);
                mine |= cov ? (1u << k) : 0u;
            }
            int base = 0;
            if (lane == 0 && cnt_before) base = (int)atomicAdd(&s_ncov, (uint32_t)cnt_before);
            base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
            for (int k = 0; k < Cfg::kPasses; k++)
                if (mine & (1u << k)) cov_list[base + my_pos[k]] = (unsigned short)(k * kScThreads + tid);
        }
        __syncthreads();
        {
            const int n_cov = (int)s_ncov;
#pragma unroll 1
            for (int i = tid; i < n_cov; i += kScThreads) {
                const int p = cov_list[i], r = p / kScW, col = p % kScW;
                const int x = g.cx0 + col, y = g.row0 + r;
                const float alpha = row_touched[r] ? covs[r * kScW + swz<true>(col)] : row_const[r];
                float4 cl = (job_paint >= 0) ? paint_at(s_paint, x, y) : make_float4(0.f, 0.f, 0.f, 0.f);
                // with_alpha: self * (alpha as f32), src/color.rs:347-349
                cl = make_float4(fmul(cl.x, alpha), fmul(cl.y, alpha), fmul(cl.z, alpha), fmul(cl.w, alpha));
                // blend_over: other + self * (1 - other.alpha), src/color.rs:342-344
                float4 d = color[p];
                const float k = fsub(1.0f, cl.w);
                d = make_float4(fadd(cl.x, fmul(d.x, k)), fadd(cl.y, fmul(d.y, k)), fadd(cl.z, fmul(d.z, k)), fadd(cl.w, fmul(d.w, k)));
                color[p] = d;
            }
        }
        dirty = true;
        __syncthreads();
    }

    // ---- the tile leaves the SM once: LinColor and / or RGBA8, coalesced rows ----
#pragma unroll 1
    for (int k = 0; k < Cfg::kPasses; k++) {
        const int p = k * kScThreads + tid, r = p / kScW, col = p % kScW;
        if (r < th && col < tw) {
            const float4 c = color[p];
            const size_t o = (size_t)(Y0 + r) * sc.width + X0 + col;
            if (dirty && sc.store_lin) sc.layer[o] = c;
            if (sc.rgba) sc.rgba[o] = lin_to_rgba8(c);
        }
    }
}

}  // namespace

// Tile width / CTA size: 256 x 8 pixels, 256 threads (4 CTAs per SM) by default — C3: 234 us against 250 us for 512 x 8 / 512
// threads and 256 us for 128 x 8 / 256; RGPU_SCENE_CW=128|256|512 and RGPU_SCENE_THREADS=256|512 select another instantiation
static int scene_env(const char* name, int dflt, int a, int b, int c) {
    const char* e = getenv(name);
    const int v = e ? atoi(e) : 0;
    return (v == a || v == b || v == c) ? v : dflt;
}
static int scene_cw() {
    static const int cw = scene_env("RGPU_SCENE_CW", 256, 128, 256, 512);
    return cw;
}
static int scene_threads() {
    static const int t = scene_env("RGPU_SCENE_THREADS", 256, 256, 512, 512);
    return (scene_cw() == 512) ? 512 : t;
}

TileShape scene_tile_shape() { return TileShape{scene_cw(), kScH}; }

template <int CW, int THREADS>
static void launch_scene_t(const JobDev* jobs, uint32_t n_jobs, const PaintDev* paints, uint32_t* tile_offs, uint32_t bin_cap,
                           const double4* bin_lines, unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, Status* status,
                           const SceneArgs& sc, bool pdl, cudaStream_t s) {
    const uint32_t n_tiles = sc.n_bands * sc.n_chunks;
    constexpr size_t smem = ScCfg<CW, THREADS>::kSmem;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(scene_kernel<CW, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(scene_kernel<CW, THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, scene_kernel<CW, THREADS>, jobs, n_jobs, paints, tile_offs, bin_cap, bin_lines, tile_state, epoch, ticket, status, sc);
}

void launch_scene(const JobDev* jobs, uint32_t n_jobs, const PaintDev* paints, uint32_t* tile_offs, uint32_t bin_cap, const double4* bin_lines,
                  unsigned long long* tile_state, uint32_t epoch, uint32_t* ticket, Status* status, const SceneArgs& sc, bool pdl,
                  cudaStream_t s) {
    if (sc.n_bands * sc.n_chunks == 0) return;
#define RGPU_SCENE(CW, T) launch_scene_t<CW, T>(jobs, n_jobs, paints, tile_offs, bin_cap, bin_lines, tile_state, epoch, ticket, status, sc, pdl, s)
    const int cw = scene_cw(), th = scene_threads();
    if (cw == 128 && th == 256) RGPU_SCENE(128, 256);
    else if (cw == 128) RGPU_SCENE(128, 512);
    else if (cw == 256 && th == 256) RGPU_SCENE(256, 256);
    else if (cw == 256) RGPU_SCENE(256, 512);
    else RGPU_SCENE(512, 512);
#undef RGPU_SCENE
}

}  // namespace rgpu
