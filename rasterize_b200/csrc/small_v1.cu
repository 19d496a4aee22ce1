// ROUND-1 VARIANT, kept for A/B (RGPU_SMALL_V1=1): see small.cu for the current kernel.
// Fused fill path for small canvases (at most 64 x 64 visible pixels, e.g. glyph batches — BASELINE config 4):
// ONE CTA per job runs K1..K4 end to end.  The path's curves are flattened straight into shared memory, the
// lines are accumulated into a shared-memory canvas, scanned, and the coverage (or the composited colour) is
// written once.  No line buffer, no bins and no carries ever touch HBM: traffic is the algorithmic minimum
// (control points in, pixels out), and a whole batch is a single launch.
//
// Same arithmetic as the tiled pipeline: flatten primitives from flatten_device.cuh (bit-exact f64, reference
// order inside a slot; order between slots is irrelevant for accumulation), span body / clipping / fixed-point
// rounding / paints from raster_device.cuh.  Reference citations live in those files.
#include "flatten_device.cuh"
#include "raster_device.cuh"

namespace rgpu {

namespace {

using namespace fl;
using namespace rs;

constexpr int kSmThreads = 256;
constexpr int kSmWarps = kSmThreads / 32;
constexpr int kSmMaxW = 64, kSmMaxH = 64;
constexpr int kSmPitch = 68;        // 64 columns + overflow column, padded to a multiple of 4 ints
constexpr int kSmLineCap = 768;     // lines kept in shared memory per window (24 KB)
constexpr int kSmRowBits = 6;
constexpr int kSmSpanCap = 256;     // per-warp span list (lanes that do not fit do their rows serially)

__global__ void __launch_bounds__(kSmThreads, 4)
small_canvas_kernel_v1(const JobDev* __restrict__ jobs, uint32_t job_first, const PaintDev* __restrict__ paints, double thr,
                    Status* __restrict__ status) {
    // dynamic shared memory (> 48 KB): line window | cells (plain row-major) | piece constants | per-warp span lists
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double4* lines_s = reinterpret_cast<double4*>(smem_raw);
    int* cells = reinterpret_cast<int*>(lines_s + kSmLineCap);
    double* p_ax = reinterpret_cast<double*>(cells + kSmMaxH * kSmPitch);
    double* p_ay = p_ax + kSmThreads;
    double* p_by = p_ay + kSmThreads;
    double* p_dxdy = p_by + kSmThreads;
    unsigned short* spans_all = reinterpret_cast<unsigned short*>(p_dxdy + kSmThreads);
    __shared__ int rowtot[kSmMaxH];
    __shared__ int row_touched[kSmMaxH];
    __shared__ uint32_t n_lines_s;
    __shared__ PaintDev s_paint;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // the job descriptor is read all through the kernel: one cooperative copy into shared memory instead of repeated
    // (L1-latency, alias-constrained) global loads
    __shared__ JobDev s_job;
    {
        const int* src = reinterpret_cast<const int*>(&jobs[job_first + blockIdx.x]);
        int* dst = reinterpret_cast<int*>(&s_job);
        if (tid < (int)(sizeof(JobDev) / 4)) dst[tid] = src[tid];
    }
    __syncthreads();
    const JobDev& job = s_job;
    const int mode = job.mode;

    {
        const int4 z = make_int4(0, 0, 0, 0);
        int4* c4 = reinterpret_cast<int4*>(cells);
        for (int i = tid; i < kSmMaxH * kSmPitch / 4; i += kSmThreads) c4[i] = z;
    }
    if (tid < kSmMaxH) { rowtot[tid] = 0; row_touched[tid] = 0; }
    const bool render = mode == kModeRender;  // fill onto a canvas created here: every pixel is written, none is read
    if (mode >= kModeFill && job.paint_index >= 0) {
        const int* src = reinterpret_cast<const int*>(&paints[job.paint_index]);
        int* dst = reinterpret_cast<int*>(&s_paint);
        for (int i = tid; i < (int)(sizeof(PaintDev) / 4); i += kSmThreads) dst[i] = src[i];
    }

    TileGeom g;
    g.row0 = 0;
    g.row1 = job.height;
    g.cx0 = 0;
    g.wc = job.clamp_w;
    g.wci = (int)g.wc;
    g.tile_end = min(kSmPitch, g.wci + 1);  // reference columns incl. the overflow column
    g.pitch = kSmPitch;
    unsigned short* spans = spans_all + warp * kSmSpanCap;
    __syncthreads();

    // ---- K1 into shared memory, K3 phase 1 from shared memory -----------------------------------------------
    // One depth-first walk per slot: every leaf claims a place in the shared line window with a shared-memory
    // atomic (the order of lines is irrelevant to the fixed-point accumulation).  If the window is full the
    // emitting thread rasterizes that line itself, so any path size works; ordinary glyphs fit in one window.
    const uint32_t total_slots = job.n_items * kSlotsPerItem;
    uint32_t my_lines = 0;
    for (uint32_t s0 = 0; s0 < total_slots; s0 += kSmThreads) {
        if (tid == 0) n_lines_s = 0;
        __syncthreads();
        // the job table entry doubles as a one-job table for slot_setup: slot t of this job is global slot
        // item_begin*8 + t of a table whose only entry starts at item_begin
        const uint32_t t = s0 + tid;
        SlotCtx c;
        if (t < total_slots && slot_setup(&job, 1, job.item_begin * kSlotsPerItem + t, thr, c, status)) {
            auto emit = [&](double x0, double y0, double x1, double y1) {
                const uint32_t k = atomicAdd(&n_lines_s, 1u);
                if (k < (uint32_t)kSmLineCap) {
                    lines_s[k] = make_double4(x0, y0, x1, y1);
                } else {  // window full: rasterize here (serial in this thread)
                    line_serial<false>(make_double4(x0, y0, x1, y1), g, cells, rowtot, row_touched);
                }
            };
            if (seg_all_finite(c.seg, c.kind)) my_lines += slot_walk<false>(c, thr, status, emit);
            else my_lines += slot_walk<true>(c, thr, status, emit);
        }
        __syncthreads();
        const uint32_t n = min(n_lines_s, (uint32_t)kSmLineCap);
        for (uint32_t i0 = warp * 32; i0 < n; i0 += kSmThreads) {
            const uint32_t i = i0 + lane;
            const bool valid = i < n;
            const double4 l = valid ? lines_s[i] : make_double4(0, 0, 0, 0);
            warp_accumulate_round<false, kSmRowBits, kSmSpanCap, unsigned short>(l, valid, g, cells, rowtot, row_touched, p_ax, p_ay, p_by, p_dxdy, spans, tid);
        }
        __syncthreads();
    }
    // line count of the batch (statistics only): one atomic per warp
    {
        uint32_t v = my_lines;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&status->n_lines, v);
    }
    __syncthreads();

    // ---- K3 phase 2 + K4: two rows per warp iteration (16 lanes x 4 columns each) ----------------------------
    const int wout = job.width_out, hout = job.height;
    const int half = lane >> 4, hl = lane & 15;
    const bool evenodd = job.rule == 1;
    // per-thread constants of the composite loop: a solid paint is its colour (glyph batches), the canvas window's base
    const bool solid = mode >= kModeFill && (job.paint_index < 0 || s_paint.kind == 0);
    const float4 solid_c = (mode >= kModeFill && job.paint_index >= 0) ? make_float4(s_paint.solid[0], s_paint.solid[1], s_paint.solid[2], s_paint.solid[3])
                                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    float4* const out_base = reinterpret_cast<float4*>(job.canvas) + job.origin;
    const unsigned long long row_stride = job.row_stride;
    for (int r2 = warp * 2; r2 < hout; r2 += kSmWarps * 2) {
        const int r = r2 + half;
        const bool rvalid = r < hout;
        int* rowc = cells + (rvalid ? r : 0) * kSmPitch;
        int4 q = make_int4(0, 0, 0, 0);
        if (rvalid) q = *reinterpret_cast<const int4*>(rowc + hl * 4);
        const int p0 = q.x, p1 = p0 + q.y, p2 = p1 + q.z, p3 = p2 + q.w;
        int incl = p3;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int nb = __shfl_up_sync(0xffffffffu, incl, o, 16);
            if (hl >= o) incl += nb;
        }
        const int base = incl - p3;
        float4 cv;
        if (evenodd)
            cv = make_float4(coverage_from_fixed<true>(base + p0), coverage_from_fixed<true>(base + p1), coverage_from_fixed<true>(base + p2),
                             coverage_from_fixed<true>(base + p3));
        else
            cv = make_float4(coverage_from_fixed<false>(base + p0), coverage_from_fixed<false>(base + p1),
                             coverage_from_fixed<false>(base + p2), coverage_from_fixed<false>(base + p3));
        const int col = hl * 4;
        if (mode < kModeFill) {
            if (rvalid) {
                if (mode == kModeCoverage) {  // mask_iter drops abs(alpha) < 1e-6 (src/rasterize.rs:348)
                    if (cv.x < 1e-6f) cv.x = 0.f;
                    if (cv.y < 1e-6f) cv.y = 0.f;
                    if (cv.z < 1e-6f) cv.z = 0.f;
                    if (cv.w < 1e-6f) cv.w = 0.f;
                }
                float* out = reinterpret_cast<float*>(job.canvas) + job.origin + (unsigned long long)r * job.row_stride;
                if (col + 3 < wout && ((reinterpret_cast<uintptr_t>(out + col) & 15) == 0)) {
                    __stcs(reinterpret_cast<float4*>(out + col), cv);
                } else {
                    if (col < wout) out[col] = cv.x;
                    if (col + 1 < wout) out[col + 1] = cv.y;
                    if (col + 2 < wout) out[col + 2] = cv.z;
                    if (col + 3 < wout) out[col + 3] = cv.w;
                }
            }
        } else {
            // stage the two rows' coverage so that consecutive lanes composite consecutive pixels (16 B each)
            if (rvalid) *reinterpret_cast<float4*>(rowc + col) = cv;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int pidx = i * 32 + lane;  // 0..127 over the two rows
                const int rr = r2 + (pidx >> 6), px = pidx & 63;
                if (rr < hout && px < wout) {
                    const float alpha = reinterpret_cast<const float*>(cells + rr * kSmPitch)[px];
                    float4* out = out_base + (unsigned long long)rr * row_stride;
                    if (alpha >= 1e-6f) {
                        float4 color = solid ? solid_c : paint_at(s_paint, px, rr);
                        color = make_float4(fmul(color.x, alpha), fmul(color.y, alpha), fmul(color.z, alpha), fmul(color.w, alpha));
                        float4 dstc = render ? make_float4(0.f, 0.f, 0.f, 0.f) : out[px];  // `Layer::new`: transparent
                        const float k = fsub(1.0f, color.w);
                        dstc = make_float4(fadd(color.x, fmul(dstc.x, k)), fadd(color.y, fmul(dstc.y, k)), fadd(color.z, fmul(dstc.z, k)),
                                           fadd(color.w, fmul(dstc.w, k)));
                        if (render) __stcs(out + px, dstc); else out[px] = dstc;
                    } else if (render) {
                        __stcs(out + px, make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace

void launch_small_canvas_v1(const JobDev* jobs, uint32_t job_first, uint32_t n_jobs, const PaintDev* paints, double thr, Status* status,
                         cudaStream_t s) {
    if (n_jobs == 0) return;
    constexpr size_t smem = sizeof(double4) * kSmLineCap + sizeof(int) * kSmMaxH * kSmPitch + sizeof(double) * 4 * kSmThreads +
                            sizeof(unsigned short) * kSmSpanCap * kSmWarps;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(small_canvas_kernel_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev] = true;
    }
    small_canvas_kernel_v1<<<n_jobs, kSmThreads, smem, s>>>(jobs, job_first, paints, thr, status);
}

}  // namespace rgpu
