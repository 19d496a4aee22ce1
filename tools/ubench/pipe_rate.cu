// Issue-rate microbenchmark: DFMA / FFMA / shared-memory integer atomics per SM per clock on the GPU at hand.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipe_rate tools/ubench/pipe_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters) {
    __shared__ int cells[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) cells[i] = 0;
    __syncthreads();
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    float f0 = (float)a0, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    unsigned h = threadIdx.x * 2654435761u;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {
            a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
        } else if (MODE == 1) {
            f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f);
        } else {
            h = h * 1664525u + 1013904223u;
            atomicAdd(&cells[(h >> 8) & 4095], 1); atomicAdd(&cells[(h >> 12) & 4095], 1);
            atomicAdd(&cells[(h >> 16) & 4095], 1); atomicAdd(&cells[(h >> 20) & 4095], 1);
        }
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + f0 + f1 + f2 + f3 + cells[threadIdx.x];
}
template <int MODE>
void run(const char* name, int sms, double mhz) {
    double* out;
    const int blocks = sms * 4, threads = 512, iters = 20000;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    k<MODE><<<blocks, threads>>>(out, 100);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * threads * iters * 4;
    printf("%s: %.1f thread-ops/clk/SM (%.3f ms, assuming %.0f MHz)\n", name, ops / (ms * 1e-3) / (mhz * 1e6) / sms, ms, mhz);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, khz);
    run<0>("DFMA", p.multiProcessorCount, khz / 1e3);
    run<1>("FFMA", p.multiProcessorCount, khz / 1e3);
    run<2>("ATOMS.ADD (random of 4096 cells)", p.multiProcessorCount, khz / 1e3);
    return 0;
}
