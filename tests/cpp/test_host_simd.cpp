// Harness of tests/test_host_simd.py: the streaming loops of rasterize_b200/csrc/host_simd.cpp against their scalar expressions.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>
namespace rgpu { size_t expand_runs_f32(const unsigned char*, size_t, size_t, const float*, float*); size_t expand_runs_f64(const unsigned char*, size_t, size_t, const float*, double*); void host_store_fence();
const char* host_simd_name(); void expand_alpha_simd(const float*, const float[4], float*, size_t); void widen_row_simd(const float*, double*, size_t); }
int main() {
    printf("simd %s\n", rgpu::host_simd_name());
    const size_t N = 100003;
    std::vector<float> a(N);
    for (size_t i = 0; i < N; i++) a[i] = (float)rand() / RAND_MAX;
    float c[4] = {0.25f, 0.5f, 0.125f, 0.75f};
    float* out = (float*)aligned_alloc(64, (N + 64) * 16);
    double* d = (double*)aligned_alloc(64, (N + 64) * 8);
    int bad = 0;
    for (size_t off = 0; off < 9; off++) for (size_t n : {0ul, 1ul, 3ul, 15ul, 16ul, 17ul, 63ul, 1000ul, N - off}) {
        memset(out, 0xff, (N + 64) * 16);
        rgpu::expand_alpha_simd(a.data() + off, c, out + 4 * off, n);
        for (size_t i = 0; i < n; i++) for (int k = 0; k < 4; k++) { float w = c[k] * a[off + i]; if (memcmp(&w, &out[4 * (off + i) + k], 4)) bad++; }
        uint32_t g; memcpy(&g, &out[4 * (off + n)], 4); if (g != 0xffffffffu) bad++;
        memset(d, 0xff, (N + 64) * 8);
        rgpu::widen_row_simd(a.data() + off, d + off, n);
        for (size_t i = 0; i < n; i++) if (d[off + i] != (double)a[off + i]) bad++;
        uint64_t h; memcpy(&h, &d[off + n], 8); if (h != ~0ull) bad++;
    }
    // run-coded rows: random classes, widths with a ragged last segment, destinations at every 4-byte / 8-byte offset class
    for (size_t width : {1ul, 63ul, 64ul, 65ul, 200ul, 4096ul, 4100ul}) for (size_t off = 0; off < 17; off += (off < 2 ? 1 : 5)) {
        const size_t segs = (width + 63) / 64;
        std::vector<unsigned char> cls(segs);
        std::vector<float> lits;
        std::vector<float> want(width);
        for (size_t s = 0; s < segs; s++) {
            cls[s] = rand() % 3;
            const size_t n = width - 64 * s < 64 ? width - 64 * s : 64;
            if (cls[s] == 2) { for (int i = 0; i < 64; i++) lits.push_back((float)rand() / RAND_MAX); }
            for (size_t i = 0; i < n; i++) want[64 * s + i] = cls[s] == 2 ? lits[lits.size() - 64 + i] : (float)cls[s];
        }
        size_t n_lit = lits.size() / 64;
        memset(out, 0xff, (N + 64) * 16);
        if (rgpu::expand_runs_f32(cls.data(), segs, width, lits.data(), out + off) != n_lit) bad++;
        rgpu::host_store_fence();
        if (memcmp(out + off, want.data(), width * 4)) bad++;
        uint32_t g; memcpy(&g, &out[off + width], 4); if (g != 0xffffffffu) bad++;
        memset(d, 0xff, (N + 64) * 8);
        if (rgpu::expand_runs_f64(cls.data(), segs, width, lits.data(), d + off) != n_lit) bad++;
        rgpu::host_store_fence();
        for (size_t i = 0; i < width; i++) if (d[off + i] != (double)want[i]) bad++;
        uint64_t h; memcpy(&h, &d[off + width], 8); if (h != ~0ull) bad++;
    }
    printf("bad %d\n", bad);
    return bad != 0;
}
