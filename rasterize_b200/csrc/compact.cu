// `Rasterizer::mask_iter` on the device (reference src/rasterize.rs:313-355): the iterator walks the (width + 1) x height
// difference image in row-major order, drops the overflow column and yields Pixel { x, y, alpha } wherever
// abs(alpha) >= 1e-6.  Here the dense coverage of RGPU_JOB_COVERAGE (0 exactly where the iterator yields nothing) is
// compacted into that list on the device, so only the yielded pixels cross PCIe (24 B each) instead of the whole canvas
// followed by a scan on one host thread:
//   pixel_count_kernel   pixels with coverage != 0 per block of 4096 consecutive pixels
//   exclusive scan       (scan.cu) -> first record of every block, total
//   pixel_emit_kernel    the records, in row-major order (ballot-free: per-lane counts of a float4, warp prefix by shuffles)
// HBM-bound: 4 B per canvas pixel read twice (the second read comes from L2 for canvases below ~100 MB) + 24 B per record.
#include "rgpu_internal.cuh"

namespace rgpu {

namespace {

constexpr int kPxThreads = 256;
constexpr int kPxWarps = kPxThreads / 32;
constexpr int kPxRounds = 4;                         // a warp takes kPxRounds x 128 consecutive pixels
constexpr int kPxPerWarp = kPxRounds * 128;
constexpr int kPxPerBlock = kPxWarps * kPxPerWarp;   // 4096

// four consecutive coverages starting at pixel i (a multiple of 4); beyond the canvas: 0
__device__ __forceinline__ float4 load4(const float* __restrict__ cov, size_t i, size_t n) {
    if (i + 4 <= n) return __ldg(reinterpret_cast<const float4*>(cov + i));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) v.x = cov[i];
    if (i + 1 < n) v.y = cov[i + 1];
    if (i + 2 < n) v.z = cov[i + 2];
    return v;
}
__device__ __forceinline__ int count4(const float4& v) { return (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f); }

__global__ void __launch_bounds__(kPxThreads) pixel_count_kernel(const float* __restrict__ cov, size_t n, uint32_t* __restrict__ counts) {
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t base = (size_t)blockIdx.x * kPxPerBlock + (size_t)warp * kPxPerWarp + 4 * lane;
    int c = 0;
#pragma unroll
    for (int r = 0; r < kPxRounds; r++) c += count4(load4(cov, base + 128 * r, n));
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c) atomicAdd(&s_total, c);
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = (uint32_t)s_total;
}

struct PixelRec {  // = rgpu_pixel (include/rasterize_b200.h): size_t x, y; double alpha
    unsigned long long x, y;
    double alpha;
};

__global__ void __launch_bounds__(kPxThreads) pixel_emit_kernel(const float* __restrict__ cov, size_t n, size_t width, const uint32_t* __restrict__ offs,
                                                                 PixelRec* __restrict__ out, size_t cap) {
    __shared__ int s_warp[kPxWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t base = (size_t)blockIdx.x * kPxPerBlock + (size_t)warp * kPxPerWarp + 4 * lane;
    float4 v[kPxRounds];
    int c[kPxRounds], mine = 0;
#pragma unroll
    for (int r = 0; r < kPxRounds; r++) {
        v[r] = load4(cov, base + 128 * r, n);
        c[r] = count4(v[r]);
        mine += c[r];
    }
    const int total = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) s_warp[warp] = total;
    __syncthreads();
    if (total == 0) return;
    size_t at = offs[blockIdx.x];
    for (int w = 0; w < warp; w++) at += s_warp[w];
#pragma unroll
    for (int r = 0; r < kPxRounds; r++) {
        // exclusive prefix of the lanes' counts in this round
        int incl = c[r];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        const int round_total = __shfl_sync(0xffffffffu, incl, 31);
        if (c[r]) {
            size_t k = at + (size_t)(incl - c[r]);
            const size_t i = base + 128 * r;
            size_t y = i / width, x = i - y * width;
            const float a[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (a[e] != 0.f) {
                    if (k < cap) {
                        out[k].x = x;
                        out[k].y = y;
                        out[k].alpha = (double)a[e];
                    }
                    k++;
                }
                if (++x == width) { x = 0; y++; }
            }
        }
        at += (size_t)round_total;
    }
}

// ---- run-coded rows: the download format of large masks -------------------------------------------------------------
// A mask is mostly constant: outside the outline 0, inside 1, fractions only along the edges.  Every row is cut into
// segments of 64 pixels; a segment whose pixels all carry the bits of +0.0f or all those of 1.0f is a class byte (0 / 1),
// any other segment (class 2) is a literal: its 64 floats, packed in (row, segment) order.  Class bytes (1/256 of the
// image), one literal offset per row and the literals cross PCIe; host threads rebuild the rows (host_simd.cpp
// expand_runs_*) — the very same bytes, but written by the CPUs' streaming stores, which outrun one PCIe link.
//   seg_classify_kernel  class byte per segment, literal count per row          (block = row)
//   exclusive scan       first literal of every row, total                      (scan.cu)
//   seg_emit_kernel      the literals, in order                                 (block = row)
constexpr int kSegPx = 64;
constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;

// the two pixels of this lane in segment s of a row; beyond the row's end: the segment's first pixel (so a ragged last
// segment is classified by its real pixels)
__device__ __forceinline__ void seg_load(const float* __restrict__ rowp, size_t width, uint32_t s, int lane, uint32_t& b0, uint32_t& b1) {
    const size_t x0 = (size_t)s * kSegPx, x = x0 + 2 * lane;
    if (x0 + kSegPx <= width && (reinterpret_cast<uintptr_t>(rowp) & 7) == 0) {  // a whole segment of an 8-byte aligned row: one load
        const uint2 v = *reinterpret_cast<const uint2*>(rowp + x);
        b0 = v.x;
        b1 = v.y;
        return;
    }
    const uint32_t first = __float_as_uint(rowp[x0]);
    b0 = x < width ? __float_as_uint(rowp[x]) : first;
    b1 = x + 1 < width ? __float_as_uint(rowp[x + 1]) : first;
}

__global__ void __launch_bounds__(kSegThreads) seg_classify_kernel(const float* __restrict__ img, size_t width, uint32_t segs, unsigned char* __restrict__ cls,
                                                                    uint32_t* __restrict__ row_cnt) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* rowp = img + (size_t)blockIdx.x * width;
    unsigned char* crow = cls + (size_t)blockIdx.x * segs;
    int cnt = 0;
    // four segments of a warp in flight at a time: the kernel waits for its loads, nothing else (ncu: long scoreboard)
    for (uint32_t s0 = warp; s0 < segs; s0 += 4 * kSegWarps) {
        uint32_t b0[4], b1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t s = s0 + u * kSegWarps;
            b0[u] = b1[u] = 0u;
            if (s < segs) seg_load(rowp, width, s, lane, b0[u], b1[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t s = s0 + u * kSegWarps;
            if (s >= segs) break;  // warp-uniform
            const uint32_t first = __shfl_sync(0xffffffffu, b0[u], 0);
            const bool same = __all_sync(0xffffffffu, b0[u] == first && b1[u] == first);
            const int c = (same && first == 0u) ? 0 : (same && first == 0x3f800000u) ? 1 : 2;
            if (lane == 0) crow[s] = (unsigned char)c;
            cnt += c == 2;
        }
    }
    if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) row_cnt[blockIdx.x] = (uint32_t)s_cnt;
}

__global__ void __launch_bounds__(kSegThreads) seg_emit_kernel(const float* __restrict__ img, size_t width, uint32_t segs, const unsigned char* __restrict__ cls,
                                                                const uint32_t* __restrict__ row_off, float* __restrict__ lits) {
    __shared__ unsigned short s_list[kSegThreads];
    __shared__ int s_warp[kSegWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* rowp = img + (size_t)blockIdx.x * width;
    const unsigned char* crow = cls + (size_t)blockIdx.x * segs;
    size_t at = row_off[blockIdx.x];
    if (row_off[blockIdx.x + 1] == at) return;  // no literal in this row
    for (uint32_t base = 0; base < segs; base += kSegThreads) {
        // the literal segments of this tile of 256 segments, in order
        const uint32_t s = base + threadIdx.x;
        const bool lit = s < segs && crow[s] == 2;
        const unsigned m = __ballot_sync(0xffffffffu, lit);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < kSegWarps; w++) {
            if (w < warp) before += s_warp[w];
            total += s_warp[w];
        }
        if (lit) s_list[before + __popc(m & ((1u << lane) - 1u))] = (unsigned short)threadIdx.x;
        __syncthreads();
        for (int k = warp; k < total; k += kSegWarps) {
            uint32_t b0, b1;
            seg_load(rowp, width, base + s_list[k], lane, b0, b1);
            reinterpret_cast<uint2*>(lits + (at + k) * kSegPx)[lane] = make_uint2(b0, b1);
        }
        at += total;
        __syncthreads();
    }
}

}  // namespace

uint32_t pixel_blocks(size_t n_pixels) { return (uint32_t)((n_pixels + kPxPerBlock - 1) / kPxPerBlock); }

void launch_pixel_count(const float* cov, size_t n_pixels, uint32_t* counts, cudaStream_t s) {
    const uint32_t nb = pixel_blocks(n_pixels);
    if (nb) pixel_count_kernel<<<nb, kPxThreads, 0, s>>>(cov, n_pixels, counts);
}

void launch_pixel_emit(const float* cov, size_t n_pixels, size_t width, const uint32_t* offs, void* out, size_t cap, cudaStream_t s) {
    const uint32_t nb = pixel_blocks(n_pixels);
    if (nb) pixel_emit_kernel<<<nb, kPxThreads, 0, s>>>(cov, n_pixels, width, offs, static_cast<PixelRec*>(out), cap);
}

uint32_t runcode_segments(size_t width) { return (uint32_t)((width + kSegPx - 1) / kSegPx); }

void launch_seg_classify(const float* img, size_t width, uint32_t rows, unsigned char* cls, uint32_t* row_cnt, cudaStream_t s) {
    if (rows && width) seg_classify_kernel<<<rows, kSegThreads, 0, s>>>(img, width, runcode_segments(width), cls, row_cnt);
}

void launch_seg_emit(const float* img, size_t width, uint32_t rows, const unsigned char* cls, const uint32_t* row_off, float* lits, cudaStream_t s) {
    if (rows && width) seg_emit_kernel<<<rows, kSegThreads, 0, s>>>(img, width, runcode_segments(width), cls, row_off, lits);
}

}  // namespace rgpu
