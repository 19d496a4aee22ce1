// Layer composition on device — the Opacity and Clip arms of `Pipeline::render_rec` (reference src/scene.rs:436-457)
// and `Layer::compose` (src/scene.rs:532-565), so that a whole scene tree renders without leaving HBM.
//
//   Clip:     child_layer.compose(mask_layer,  |dst, src| dst * (src as f32))        -> scale_by_mask_kernel
//             layer.compose(child_layer,       |dst, src| dst.blend_over(src))       -> blend_over_kernel (opacity unused)
//   Opacity:  layer.compose(child_layer,       |dst, src| dst.blend_over(src * o))   -> blend_over_kernel
//
// Both are pure streaming kernels over the intersection rectangle of the two layers (computed by the host exactly as
// `Layer::compose` does): 16 B LinColor pixels, one thread per pixel, rows walked by blockIdx.y so that consecutive
// threads touch consecutive pixels.  Colour maths is f32 and unfused like the reference's SSE code
// (`LinColor` Mul<f32> and `blend_over`, src/color.rs:342-349).
#include "raster_device.cuh"

#include <algorithm>

namespace rgpu {

namespace {

using namespace rs;

__global__ void __launch_bounds__(256)
scale_by_mask_kernel(float4* __restrict__ lin, unsigned long long lin_stride, const float* __restrict__ mask, unsigned long long mask_stride,
                     uint32_t width, uint32_t height) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (uint32_t y = blockIdx.y; y < height; y += gridDim.y) {
        const float m = mask[(unsigned long long)y * mask_stride + x];
        float4* p = lin + (unsigned long long)y * lin_stride + x;
        const float4 c = *p;
        *p = make_float4(fmul(c.x, m), fmul(c.y, m), fmul(c.z, m), fmul(c.w, m));
    }
}

template <bool OPACITY>
__global__ void __launch_bounds__(256)
blend_over_kernel(float4* __restrict__ dst, unsigned long long dst_stride, const float4* __restrict__ src, unsigned long long src_stride,
                  uint32_t width, uint32_t height, float opacity) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (uint32_t y = blockIdx.y; y < height; y += gridDim.y) {
        float4 s = src[(unsigned long long)y * src_stride + x];
        if (OPACITY) s = make_float4(fmul(s.x, opacity), fmul(s.y, opacity), fmul(s.z, opacity), fmul(s.w, opacity));
        float4* p = dst + (unsigned long long)y * dst_stride + x;
        const float4 d = *p;
        // blend_over: other + self * (1 - other.alpha), src/color.rs:342-344
        const float k = fsub(1.0f, s.w);
        *p = make_float4(fadd(s.x, fmul(d.x, k)), fadd(s.y, fmul(d.y, k)), fadd(s.z, fmul(d.z, k)), fadd(s.w, fmul(d.w, k)));
    }
}

dim3 rect_grid(uint32_t width, uint32_t height) {
    const uint32_t gx = (width + 255) / 256;
    // enough rows in flight to fill 148 SMs several times over without launching one CTA per row of a huge layer
    const uint32_t gy = std::min<uint32_t>(height, std::max<uint32_t>(1u, (148u * 8u + gx - 1) / gx));
    return dim3(gx, gy);
}

}  // namespace

void launch_scale_by_mask(float4* lin, unsigned long long lin_stride, const float* mask, unsigned long long mask_stride, uint32_t width,
                          uint32_t height, cudaStream_t s) {
    if (width == 0 || height == 0) return;
    scale_by_mask_kernel<<<rect_grid(width, height), 256, 0, s>>>(lin, lin_stride, mask, mask_stride, width, height);
}

void launch_blend_over(float4* dst, unsigned long long dst_stride, const float4* src, unsigned long long src_stride, uint32_t width,
                       uint32_t height, bool use_opacity, float opacity, cudaStream_t s) {
    if (width == 0 || height == 0) return;
    if (use_opacity) blend_over_kernel<true><<<rect_grid(width, height), 256, 0, s>>>(dst, dst_stride, src, src_stride, width, height, opacity);
    else blend_over_kernel<false><<<rect_grid(width, height), 256, 0, s>>>(dst, dst_stride, src, src_stride, width, height, 1.0f);
}

}  // namespace rgpu
