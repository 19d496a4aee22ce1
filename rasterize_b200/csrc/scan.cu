// Exclusive prefix sum (u32) — single pass, decoupled look-back.
//
// Used twice per batch: over the per-slot line counts of K1 (to place every slot's lines in the reference's
// order) and over the per-band reference counts of K2.  One CTA scans a tile of 2048 values; tiles publish
// {flag, value} packed in one 64-bit word so a predecessor's aggregate / inclusive prefix is observed atomically.
// Tile ids are handed out by an atomic counter so a tile never waits on a CTA that has not started.
#include "rgpu_internal.cuh"

namespace rgpu {

namespace {

constexpr int kScanThreads = 256;
constexpr int kItemsPerThread = 8;
constexpr int kTile = kScanThreads * kItemsPerThread;
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kFlagMask = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, unsigned long long* __restrict__ state,
            uint32_t* __restrict__ counter) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * kTile + tid * kItemsPerThread;

    uint32_t v[kItemsPerThread];
    if (base + kItemsPerThread <= n) {
        const uint4 a = *reinterpret_cast<const uint4*>(in + base);
        const uint4 b = *reinterpret_cast<const uint4*>(in + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < kItemsPerThread; i++) v[i] = (base + i < n) ? in[base + i] : 0u;
    }
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < kItemsPerThread; i++) tsum += v[i];
    // block-wide exclusive scan of the thread sums
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nb;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        uint32_t s = s_warp[w];
        if (w < warp) warp_off += s;
        block_sum += s;
    }
    uint32_t thread_excl = warp_off + incl - tsum;

    if (warp == 0) {
        if (lane == 0) st_state(&state[tile], (tile == 0 ? kFlagPrefix : kFlagAgg) | block_sum);
        uint32_t excl = 0;
        if (tile > 0) {
            int look = (int)tile - 1;
            while (true) {
                int idx = look - lane;
                unsigned long long st = idx >= 0 ? ld_state(&state[idx]) : kFlagPrefix;
                while (__any_sync(0xffffffffu, (st & kFlagMask) == 0ull)) st = idx >= 0 ? ld_state(&state[idx]) : kFlagPrefix;
                unsigned pm = __ballot_sync(0xffffffffu, (st & kFlagMask) == kFlagPrefix);
                int first = pm ? (__ffs(pm) - 1) : 32;
                uint32_t val = (lane <= first) ? (uint32_t)(st & 0xffffffffull) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                excl += val;
                if (pm) break;
                look -= 32;
            }
            if (lane == 0) st_state(&state[tile], kFlagPrefix | (unsigned long long)(excl + block_sum));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    uint32_t run = s_prefix + thread_excl;
    if (base + kItemsPerThread <= n) {
        uint32_t o[kItemsPerThread];
#pragma unroll
        for (int i = 0; i < kItemsPerThread; i++) { o[i] = run; run += v[i]; }
        *reinterpret_cast<uint4*>(out + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(out + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int i = 0; i < kItemsPerThread; i++) {
            if (base + i < n) out[base + i] = run;
            run += v[i];
        }
    }
}

}  // namespace

size_t scan_temp_bytes(uint32_t n) {
    size_t tiles = ((size_t)n + kTile - 1) / kTile;
    return 16 + tiles * sizeof(unsigned long long);
}

// `temp` layout: [0..4) tile counter, [16..) tile states
void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, void* temp, size_t temp_bytes, cudaStream_t s, bool temp_is_zero) {
    if (n == 0) return;
    uint32_t tiles = (n + kTile - 1) / kTile;
    (void)temp_bytes;
    if (!temp_is_zero) cudaMemsetAsync(temp, 0, 16 + (size_t)tiles * sizeof(unsigned long long), s);
    auto* counter = static_cast<uint32_t*>(temp);
    auto* state = reinterpret_cast<unsigned long long*>(static_cast<char*>(temp) + 16);
    scan_kernel<<<tiles, kScanThreads, 0, s>>>(in, out, n, state, counter);
}

}  // namespace rgpu
