// Host-side streaming loops of the host-buffer entry points, with the widest vector unit the CPU has (picked once at run
// time): the destination is the caller's image, written once and not read here, so every variant uses non-temporal stores —
// whole cache lines at a time with AVX-512.  Plumbing only: one IEEE single-precision multiplication per component
// (expand_alpha: exactly the `colour * alpha` the glyph kernel does for a plain solid paint) or an exact f32 -> f64
// conversion (widen_row); no rasterization happens here.  Compiled by the host compiler (target attributes per function).
#include <cstddef>
#include <cstdint>
#include <immintrin.h>

namespace rgpu {

namespace {

void expand_scalar(const float* alpha, const float c[4], float* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        const float a = alpha[i];
        out[4 * i] = c[0] * a;
        out[4 * i + 1] = c[1] * a;
        out[4 * i + 2] = c[2] * a;
        out[4 * i + 3] = c[3] * a;
    }
}

void expand_sse2(const float* alpha, const float c[4], float* out, size_t n) {
    const __m128 cv = _mm_loadu_ps(c);
    for (size_t i = 0; i < n; i++) _mm_stream_ps(out + 4 * i, _mm_mul_ps(cv, _mm_set1_ps(alpha[i])));
}

__attribute__((target("avx2"))) void expand_avx2(const float* alpha, const float c[4], float* out, size_t n) {
    const __m256 cv = _mm256_broadcast_ps(reinterpret_cast<const __m128*>(c));
    const __m256i lo = _mm256_setr_epi32(0, 0, 0, 0, 1, 1, 1, 1), hi = _mm256_setr_epi32(2, 2, 2, 2, 3, 3, 3, 3);
    const __m256i lo2 = _mm256_setr_epi32(4, 4, 4, 4, 5, 5, 5, 5), hi2 = _mm256_setr_epi32(6, 6, 6, 6, 7, 7, 7, 7);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m256 a = _mm256_loadu_ps(alpha + i);
        _mm256_stream_ps(out + 4 * i, _mm256_mul_ps(cv, _mm256_permutevar8x32_ps(a, lo)));
        _mm256_stream_ps(out + 4 * i + 8, _mm256_mul_ps(cv, _mm256_permutevar8x32_ps(a, hi)));
        _mm256_stream_ps(out + 4 * i + 16, _mm256_mul_ps(cv, _mm256_permutevar8x32_ps(a, lo2)));
        _mm256_stream_ps(out + 4 * i + 24, _mm256_mul_ps(cv, _mm256_permutevar8x32_ps(a, hi2)));
    }
    const __m128 c4 = _mm_loadu_ps(c);
    for (; i < n; i++) _mm_stream_ps(out + 4 * i, _mm_mul_ps(c4, _mm_set1_ps(alpha[i])));
}

__attribute__((target("avx512f"))) void expand_avx512(const float* alpha, const float c[4], float* out, size_t n) {
    const __m512 cv = _mm512_broadcast_f32x4(_mm_loadu_ps(c));
    const __m512i i0 = _mm512_setr_epi32(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3);
    const __m512i four = _mm512_set1_epi32(4);
    const __m512i i1 = _mm512_add_epi32(i0, four), i2 = _mm512_add_epi32(i1, four), i3 = _mm512_add_epi32(i2, four);
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m512 a = _mm512_loadu_ps(alpha + i);
        _mm512_stream_ps(out + 4 * i, _mm512_mul_ps(cv, _mm512_permutexvar_ps(i0, a)));
        _mm512_stream_ps(out + 4 * i + 16, _mm512_mul_ps(cv, _mm512_permutexvar_ps(i1, a)));
        _mm512_stream_ps(out + 4 * i + 32, _mm512_mul_ps(cv, _mm512_permutexvar_ps(i2, a)));
        _mm512_stream_ps(out + 4 * i + 48, _mm512_mul_ps(cv, _mm512_permutexvar_ps(i3, a)));
    }
    const __m128 c4 = _mm_loadu_ps(c);
    for (; i < n; i++) _mm_stream_ps(out + 4 * i, _mm_mul_ps(c4, _mm_set1_ps(alpha[i])));
}

void widen_sse2(const float* src, double* dst, size_t n) {
    size_t x = 0;
    for (; x + 4 <= n; x += 4) {
        const __m128 v = _mm_loadu_ps(src + x);
        _mm_stream_pd(dst + x, _mm_cvtps_pd(v));
        _mm_stream_pd(dst + x + 2, _mm_cvtps_pd(_mm_movehl_ps(v, v)));
    }
    for (; x < n; x++) dst[x] = (double)src[x];
}

__attribute__((target("avx512f"))) void widen_avx512(const float* src, double* dst, size_t n) {
    size_t x = 0;
    for (; x + 16 <= n; x += 16) {
        _mm512_stream_pd(dst + x, _mm512_cvtps_pd(_mm256_loadu_ps(src + x)));
        _mm512_stream_pd(dst + x + 8, _mm512_cvtps_pd(_mm256_loadu_ps(src + x + 8)));
    }
    for (; x + 2 <= n; x += 2) _mm_stream_pd(dst + x, _mm_cvtps_pd(_mm_castsi128_ps(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(src + x)))));
    for (; x < n; x++) dst[x] = (double)src[x];
}

// ---- run-coded rows (compact.cu): one row of class bytes + literals -> pixels ----
// a segment of n <= 64 pixels: the constant v, or the literal's floats
template <class T>
inline void seg_portable(T* d, const float* lit, T v, size_t n) {
    if (lit) for (size_t i = 0; i < n; i++) d[i] = (T)lit[i];
    else for (size_t i = 0; i < n; i++) d[i] = v;
}

__attribute__((target("avx512f"))) size_t runs_f32_avx512(const unsigned char* cls, size_t n_segs, size_t width, const float* lits, float* dst) {
    const __m512 zero = _mm512_setzero_ps(), one = _mm512_set1_ps(1.0f);
    size_t lit = 0;
    for (size_t s = 0; s < n_segs; s++) {
        float* d = dst + 64 * s;
        const size_t n = width - 64 * s < 64 ? width - 64 * s : 64;
        const unsigned c = cls[s];
        const float* src = c == 2 ? lits + 64 * lit++ : nullptr;
        if (n == 64 && (reinterpret_cast<uintptr_t>(d) & 63) == 0) {
            if (src) {
                _mm512_stream_ps(d, _mm512_loadu_ps(src));
                _mm512_stream_ps(d + 16, _mm512_loadu_ps(src + 16));
                _mm512_stream_ps(d + 32, _mm512_loadu_ps(src + 32));
                _mm512_stream_ps(d + 48, _mm512_loadu_ps(src + 48));
            } else {
                const __m512 v = c ? one : zero;
                _mm512_stream_ps(d, v);
                _mm512_stream_ps(d + 16, v);
                _mm512_stream_ps(d + 32, v);
                _mm512_stream_ps(d + 48, v);
            }
        } else {
            seg_portable<float>(d, src, c ? 1.0f : 0.0f, n);
        }
    }
    return lit;
}

size_t runs_f32_sse2(const unsigned char* cls, size_t n_segs, size_t width, const float* lits, float* dst) {
    size_t lit = 0;
    for (size_t s = 0; s < n_segs; s++) {
        float* d = dst + 64 * s;
        const size_t n = width - 64 * s < 64 ? width - 64 * s : 64;
        const unsigned c = cls[s];
        const float* src = c == 2 ? lits + 64 * lit++ : nullptr;
        if (n == 64 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
            const __m128 v = _mm_set1_ps(c ? 1.0f : 0.0f);
            for (int i = 0; i < 64; i += 4) _mm_stream_ps(d + i, src ? _mm_loadu_ps(src + i) : v);
        } else {
            seg_portable<float>(d, src, c ? 1.0f : 0.0f, n);
        }
    }
    return lit;
}

__attribute__((target("avx512f"))) size_t runs_f64_avx512(const unsigned char* cls, size_t n_segs, size_t width, const float* lits, double* dst) {
    const __m512d zero = _mm512_setzero_pd(), one = _mm512_set1_pd(1.0);
    size_t lit = 0;
    for (size_t s = 0; s < n_segs; s++) {
        double* d = dst + 64 * s;
        const size_t n = width - 64 * s < 64 ? width - 64 * s : 64;
        const unsigned c = cls[s];
        const float* src = c == 2 ? lits + 64 * lit++ : nullptr;
        if (n == 64 && (reinterpret_cast<uintptr_t>(d) & 63) == 0) {
            if (src) {
                for (int i = 0; i < 64; i += 8) _mm512_stream_pd(d + i, _mm512_cvtps_pd(_mm256_loadu_ps(src + i)));
            } else {
                const __m512d v = c ? one : zero;
                for (int i = 0; i < 64; i += 8) _mm512_stream_pd(d + i, v);
            }
        } else {
            seg_portable<double>(d, src, c ? 1.0 : 0.0, n);
        }
    }
    return lit;
}

size_t runs_f64_sse2(const unsigned char* cls, size_t n_segs, size_t width, const float* lits, double* dst) {
    size_t lit = 0;
    for (size_t s = 0; s < n_segs; s++) {
        double* d = dst + 64 * s;
        const size_t n = width - 64 * s < 64 ? width - 64 * s : 64;
        const unsigned c = cls[s];
        const float* src = c == 2 ? lits + 64 * lit++ : nullptr;
        if (n == 64 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
            const __m128d v = _mm_set1_pd(c ? 1.0 : 0.0);
            for (int i = 0; i < 64; i += 2)
                _mm_stream_pd(d + i, src ? _mm_cvtps_pd(_mm_castsi128_ps(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(src + i)))) : v);
        } else {
            seg_portable<double>(d, src, c ? 1.0 : 0.0, n);
        }
    }
    return lit;
}

int simd_level() {  // 0: SSE2, 1: AVX2, 2: AVX-512F
    static const int level = [] {
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f")) return 2;
        if (__builtin_cpu_supports("avx2")) return 1;
        return 0;
    }();
    return level;
}

}  // namespace

const char* host_simd_name() { return simd_level() == 2 ? "avx512f" : simd_level() == 1 ? "avx2" : "sse2"; }

// out[4 i + k] = colour[k] * alpha[i], streaming stores
void expand_alpha_simd(const float* alpha, const float colour[4], float* out, size_t n) {
    if (reinterpret_cast<uintptr_t>(out) & 15) return expand_scalar(alpha, colour, out, n);
    // up to the next 64-byte boundary of the destination one pixel (16 B) at a time
    size_t head = 0;
    while (head < n && (reinterpret_cast<uintptr_t>(out + 4 * head) & 63)) head++;
    expand_sse2(alpha, colour, out, head);
    alpha += head;
    out += 4 * head;
    n -= head;
    switch (simd_level()) {
        case 2: expand_avx512(alpha, colour, out, n); break;
        case 1: expand_avx2(alpha, colour, out, n); break;
        default: expand_sse2(alpha, colour, out, n); break;
    }
    _mm_sfence();
}

// dst[i] = (double)src[i], streaming stores
void widen_row_simd(const float* src, double* dst, size_t n) {
    size_t x = 0;
    while (x < n && (reinterpret_cast<uintptr_t>(dst + x) & 63)) {
        dst[x] = (double)src[x];
        x++;
    }
    if (simd_level() == 2) widen_avx512(src + x, dst + x, n - x);
    else widen_sse2(src + x, dst + x, n - x);
}

// One row of a run-coded image (compact.cu): class byte per 64-pixel segment (0: zeros, 1: ones, 2: the next literal of
// `lits`, 64 floats each) -> `width` pixels at dst.  Returns the literals consumed.  The caller fences (host_store_fence).
size_t expand_runs_f32(const unsigned char* cls, size_t n_segs, size_t width, const float* lits, float* dst) {
    return simd_level() == 2 ? runs_f32_avx512(cls, n_segs, width, lits, dst) : runs_f32_sse2(cls, n_segs, width, lits, dst);
}
size_t expand_runs_f64(const unsigned char* cls, size_t n_segs, size_t width, const float* lits, double* dst) {
    return simd_level() == 2 ? runs_f64_avx512(cls, n_segs, width, lits, dst) : runs_f64_sse2(cls, n_segs, width, lits, dst);
}
void host_store_fence() { _mm_sfence(); }

}  // namespace rgpu
