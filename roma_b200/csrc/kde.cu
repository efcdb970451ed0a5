// sample(): Gaussian kernel density of the 4-D match coordinates (romatch/utils/kde.py:4-12) as a tiled
// all-pairs reduction that never materialises the N x N matrix (the reference builds a 40000^2 fp16
// matrix = 3.2 GB).  half != 0 reproduces the reference's fp16 arithmetic: inputs, the distance, its
// square/scale and the exponential are rounded to fp16 step by step; the row sum accumulates in fp32
// and is rounded to fp16 at the end, as torch's half `sum` does.
#include "common.cuh"

namespace rb {

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }

__global__ void __launch_bounds__(128) kde_kernel(const float* __restrict__ x, float* __restrict__ density, int n, float inv_two_var, int half) {
    __shared__ float4 pts[512];
    const int i = blockIdx.x * 128 + threadIdx.x;
    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) xi = reinterpret_cast<const float4*>(x)[i];
    if (half) { xi.x = rh(xi.x); xi.y = rh(xi.y); xi.z = rh(xi.z); xi.w = rh(xi.w); }
    float acc = 0.f;
    for (int j0 = 0; j0 < n; j0 += 512) {
        for (int t = threadIdx.x; t < 512; t += 128) {
            float4 v = make_float4(1e30f, 1e30f, 1e30f, 1e30f);
            if (j0 + t < n) {
                v = reinterpret_cast<const float4*>(x)[j0 + t];
                if (half) { v.x = rh(v.x); v.y = rh(v.y); v.z = rh(v.z); v.w = rh(v.w); }
            }
            pts[t] = v;
        }
        __syncthreads();
        int lim = min(512, n - j0);
        for (int t = 0; t < lim; ++t) {
            float4 v = pts[t];
            float dx = xi.x - v.x, dy = xi.y - v.y, dz = xi.z - v.z, dw = xi.w - v.w;
            float d2 = dx * dx + dy * dy + dz * dz + dw * dw;
            float e;
            if (half) {
                float d = rh(sqrtf(d2));            // cdist result in fp16
                float q = rh(rh(-rh(d * d)) * inv_two_var);    // (-d**2) / (2 std^2)
                e = rh(expf(q));
            } else {
                e = expf(-d2 * inv_two_var);
            }
            acc += e;
        }
        __syncthreads();
    }
    if (i < n) density[i] = half ? rh(acc) : acc;
}

}  // namespace rb

extern "C" int romab200_kde_density(const rb_kde_args* a, void* stream) {
    using namespace rb;
    cudaStream_t st = (cudaStream_t)stream;
    RB_REQUIRE(a->n > 0 && ((uintptr_t)a->x) % 16 == 0, "kde_density: n=%d or unaligned input", a->n);
    float inv = 1.0f / (2.0f * a->std * a->std);
    kde_kernel<<<(a->n + 127) / 128, 128, 0, st>>>(a->x, a->density, a->n, inv, a->half);
    return check_launch("kde_density");
}
