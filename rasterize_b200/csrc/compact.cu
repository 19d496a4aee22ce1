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

}  // namespace rgpu
