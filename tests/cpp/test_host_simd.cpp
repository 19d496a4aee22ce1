// Harness of tests/test_host_simd.py: the streaming loops of rasterize_b200/csrc/host_simd.cpp against their scalar expressions.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>
namespace rgpu { const char* host_simd_name(); void expand_alpha_simd(const float*, const float[4], float*, size_t); void widen_row_simd(const float*, double*, size_t); }
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
    printf("bad %d\n", bad);
    return bad != 0;
}
