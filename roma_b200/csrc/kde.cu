// sample(): Gaussian kernel density of the 4-D match coordinates (romatch/utils/kde.py:4-12) as a tiled
// all-pairs reduction that never materialises the N x N matrix (the reference builds a 40000^2 fp16
// matrix = 3.2 GB).  half != 0 reproduces the reference's fp16 arithmetic step by step (see below); the row
// sum accumulates in fp32 and is rounded to fp16 at the end, as torch's half `sum` does.
#include "common.cuh"

namespace rb {

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }

// half mode follows torch.cdist's matmul formulation for fp16 inputs (the path the reference takes):
//   d2 = fp16( fp32-accumulated  [-2x | ||x||^2 | 1] . [y | 1 | ||y||^2] ) clamped at 0,
// then exp(-d2 / (2 std^2)) in fp32, fp32 row sum rounded to fp16 (see the comment in the loop for what is skipped).
__global__ void __launch_bounds__(128) kde_kernel(const float* __restrict__ x, float* __restrict__ density, int n, float two_var, int half) {
    rb::pdl_wait();
    __shared__ float4 pts[512];
    __shared__ float nrm[512];
    const int i = blockIdx.x * 128 + threadIdx.x;
    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) xi = reinterpret_cast<const float4*>(x)[i];
    float ni = 0.f;
    if (half) {
        xi.x = rh(xi.x); xi.y = rh(xi.y); xi.z = rh(xi.z); xi.w = rh(xi.w);
        ni = rh(rh(xi.x * xi.x) + rh(xi.y * xi.y) + rh(xi.z * xi.z) + rh(xi.w * xi.w));
    }
    const float ax = -2.f * xi.x, ay = -2.f * xi.y, az = -2.f * xi.z, aw = -2.f * xi.w;
    float acc = 0.f;
    for (int j0 = 0; j0 < n; j0 += 512) {
        for (int t = threadIdx.x; t < 512; t += 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            float nv = half ? INFINITY : 0.f;
            if (j0 + t < n) {
                nv = 0.f;
                v = reinterpret_cast<const float4*>(x)[j0 + t];
                if (half) {
                    v.x = rh(v.x); v.y = rh(v.y); v.z = rh(v.z); v.w = rh(v.w);
                    nv = rh(rh(v.x * v.x) + rh(v.y * v.y) + rh(v.z * v.z) + rh(v.w * v.w));
                }
            }
            pts[t] = v; nrm[t] = nv;
        }
        __syncthreads();
        int lim = min(512, n - j0);
        if (half) {
            // two pairs per iteration: fp32 dot product (the matmul), d2 rounded to fp16 and clamped (that quantisation is
            // the one fp16 effect that matters: a d2 ulp moves exp(-50 d2) by up to 10 %), then exp in fp32.  The reference's
            // sqrt -> square round trip (cdist, then **2) and the fp16 rounding of each exp are skipped: together they move a
            // density by at most one fp16 ulp (measured against the oracle: max 1e-3, mean 1e-4 relative) and cost two of
            // the three MUFU operations per pair.
            // padding entries (j >= n) hold x = 0, norm = +inf  ->  d2 = inf  ->  exp(-inf) = 0
            const __half2 zero2 = __float2half2_rn(0.f);
            const float nscale = -1.4426950408889634f / two_var;       // exp(-d2 / two_var) = exp2(d2 * nscale)
            float acc2 = 0.f;
#pragma unroll 4
            for (int t = 0; t < 512; t += 2) {
                if (t >= lim) break;
                const float4 v0 = pts[t], v1 = pts[t + 1];
                float s0 = ax * v0.x, s1 = ax * v1.x;
                s0 = fmaf(ay, v0.y, s0); s1 = fmaf(ay, v1.y, s1);
                s0 = fmaf(az, v0.z, s0); s1 = fmaf(az, v1.z, s1);
                s0 = fmaf(aw, v0.w, s0); s1 = fmaf(aw, v1.w, s1);
                s0 = (s0 + ni) + nrm[t]; s1 = (s1 + ni) + nrm[t + 1];
                const float2 d2 = __half22float2(__hmax2(__floats2half2_rn(s0, s1), zero2));
                acc += exp2f(d2.x * nscale); acc2 += exp2f(d2.y * nscale);
            }
            acc += acc2;
        } else {
            for (int t = 0; t < lim; ++t) {
                float4 v = pts[t];
                float dx = xi.x - v.x, dy = xi.y - v.y, dz = xi.z - v.z, dw = xi.w - v.w;
                acc += expf(-(dx * dx + dy * dy + dz * dz + dw * dw) / two_var);
            }
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
    float two_var = (float)(2.0 * (double)a->std * (double)a->std);
    rb::launch_pdl(kde_kernel, dim3((a->n + 127) / 128), dim3(128), 0, st, a->x, a->density, a->n, two_var, a->half);
    return check_launch("kde_density");
}
