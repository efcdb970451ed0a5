// Image preprocessing in front of match(): RGB uint8 image -> network input, on the device.
// The reference does this on the host with get_tuple_transform_ops(resize=(h, w), normalize=True) (romatch/utils/utils.py:164-173,
// called at romatch/models/matcher.py:812-815,855-866): torchvision Resize on a PIL image = PIL.Image.resize((w, h), BICUBIC)
// (utils.py:233-238), np.array(float32) / 255 (utils.py:175-183), ImageNet mean / std (utils.py:250-260).
//
// Pillow's 8-bit resampling is integer arithmetic, so the device result is the same bytes: per output coordinate a window of input samples
// with 22-bit fixed-point weights (bicubic a = -0.5, support 2 * max(1, in/out), normalised in double precision, rounded half away from
// zero), horizontal pass into a uint8 image, then the vertical pass; each pass is clip8((2^21 + sum(pixel * k)) >> 22).  The weight tables
// depend on (in, out) only: romab200_resample_coeffs builds them on the HOST, the caller caches a device copy per size pair.
// The float stage keeps the reference's operation order (divide by 255, subtract mean, divide by std, all IEEE fp32).
#include "common.cuh"
#include <cmath>
#include <vector>

namespace rb {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

static inline double bicubic_weight(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[y, xx, c] = clip8(2^21 + sum_x in[y, xmin + x, c] * k[xx][x])
__global__ void __launch_bounds__(128) resample_rows_kernel(const uint8_t* __restrict__ in, int64_t ld_in, int in_h, uint8_t* __restrict__ out,
                                                            int out_w, const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (xx >= out_w || y >= in_h) return;
    const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const int32_t* k = kk + (int64_t)xx * ksize;
    const uint8_t* src = in + (int64_t)y * ld_in + (int64_t)xmin * 3;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < cnt; ++x) {
        const int w = k[x];
        s0 += (int)src[3 * x] * w; s1 += (int)src[3 * x + 1] * w; s2 += (int)src[3 * x + 2] * w;
    }
    uint8_t* dst = out + ((int64_t)y * out_w + xx) * 3;
    dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

struct PreprocessTail {
    const uint8_t* src; int64_t ld_src; int src_h;       // the horizontally resampled image (or the input when the width is unchanged)
    int out_h, out_w;
    const int32_t* bounds; const int32_t* kk; int ksize;   // vertical tables, bounds == nullptr: the height is unchanged
    uint8_t* out_u8; float* out;
    float mean[3], stdv[3];
};

// vertical pass + /255 + mean/std; one thread per output pixel
__global__ void __launch_bounds__(128) resample_cols_normalize_kernel(const PreprocessTail p) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    if (xx >= p.out_w || yy >= p.out_h) return;
    uint8_t px[3];
    if (p.bounds) {
        const int ymin = p.bounds[2 * yy], cnt = p.bounds[2 * yy + 1];
        const int32_t* k = p.kk + (int64_t)yy * p.ksize;
        const uint8_t* src = p.src + (int64_t)ymin * p.ld_src + (int64_t)xx * 3;
        int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < cnt; ++y) {
            const int w = k[y];
            const uint8_t* r = src + (int64_t)y * p.ld_src;
            s0 += (int)r[0] * w; s1 += (int)r[1] * w; s2 += (int)r[2] * w;
        }
        px[0] = clip8(s0); px[1] = clip8(s1); px[2] = clip8(s2);
    } else {
        const uint8_t* r = p.src + (int64_t)yy * p.ld_src + (int64_t)xx * 3;
        px[0] = r[0]; px[1] = r[1]; px[2] = r[2];
    }
    if (p.out_u8) {
        uint8_t* d = p.out_u8 + ((int64_t)yy * p.out_w + xx) * 3;
        d[0] = px[0]; d[1] = px[1]; d[2] = px[2];
    }
    if (p.out) {
        const int64_t plane = (int64_t)p.out_h * p.out_w, o = (int64_t)yy * p.out_w + xx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fdiv_rn((float)px[c], 255.0f);
            p.out[c * plane + o] = __fdiv_rn(__fsub_rn(v, p.mean[c]), p.stdv[c]);
        }
    }
}

}  // namespace rb

using namespace rb;

// Host only (no CUDA call): Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter over the whole axis.
extern "C" int romab200_resample_coeffs(const rb_resample_coeffs_args* a, void* /*stream*/) {
    {
        RB_REQUIRE(a && a->in_size > 0 && a->out_size > 0 && a->ksize, "resample_coeffs: bad sizes");
        const double scale = (double)a->in_size / (double)a->out_size;
        double filterscale = scale;
        if (filterscale < 1.0) filterscale = 1.0;
        const double support = 2.0 * filterscale;
        const int ksize = (int)std::ceil(support) * 2 + 1;
        *a->ksize = ksize;
        if (!a->kk && !a->bounds) return 0;
        RB_REQUIRE(a->kk && a->bounds, "resample_coeffs: kk and bounds come together");
        const double ss = 1.0 / filterscale;
        std::vector<double> w(ksize);
        for (int xx = 0; xx < a->out_size; ++xx) {
            const double center = (xx + 0.5) * scale;
            int xmin = (int)(center - support + 0.5);
            if (xmin < 0) xmin = 0;
            int xmax = (int)(center + support + 0.5);
            if (xmax > a->in_size) xmax = a->in_size;
            xmax -= xmin;
            double ww = 0.0;
            for (int x = 0; x < xmax; ++x) { w[x] = bicubic_weight((x + xmin - center + 0.5) * ss); ww += w[x]; }
            int32_t* k = a->kk + (int64_t)xx * ksize;
            for (int x = 0; x < ksize; ++x) {
                if (x >= xmax) { k[x] = 0; continue; }
                const double v = ww != 0.0 ? w[x] / ww : w[x];
                k[x] = v < 0 ? (int32_t)(-0.5 + v * (double)(1 << RS_PRECISION_BITS)) : (int32_t)(0.5 + v * (double)(1 << RS_PRECISION_BITS));
            }
            a->bounds[2 * xx] = xmin; a->bounds[2 * xx + 1] = xmax;
        }
        return 0;
    }
}

extern "C" int romab200_preprocess_rgb8(const rb_preprocess_args* a, void* stream) {
    {
        RB_REQUIRE(a && a->in && a->in_h > 0 && a->in_w > 0 && a->out_h > 0 && a->out_w > 0, "preprocess_rgb8: bad sizes");
        RB_REQUIRE(a->ld_in >= (int64_t)a->in_w * 3, "preprocess_rgb8: row pitch smaller than 3 * width");
        RB_REQUIRE(a->out || a->out_u8, "preprocess_rgb8: no output");
        RB_REQUIRE(a->in_h <= 65535 && a->out_h <= 65535, "preprocess_rgb8: more than 65535 rows");
        const bool need_x = a->out_w != a->in_w, need_y = a->out_h != a->in_h;
        RB_REQUIRE(!need_x || (a->bounds_x && a->kk_x && a->ksize_x > 0 && a->tmp), "preprocess_rgb8: horizontal tables / tmp missing");
        RB_REQUIRE(!need_y || (a->bounds_y && a->kk_y && a->ksize_y > 0), "preprocess_rgb8: vertical tables missing");
        cudaStream_t st = (cudaStream_t)stream;
        if (need_x) {
            dim3 grid((a->out_w + 127) / 128, a->in_h);
            resample_rows_kernel<<<grid, 128, 0, st>>>(a->in, a->ld_in, a->in_h, a->tmp, a->out_w, a->bounds_x, a->kk_x, a->ksize_x);
            if (int rc = check_launch("resample_rows")) return rc;
        }
        PreprocessTail p;
        p.src = need_x ? a->tmp : a->in; p.ld_src = need_x ? (int64_t)a->out_w * 3 : a->ld_in; p.src_h = a->in_h;
        p.out_h = a->out_h; p.out_w = a->out_w;
        p.bounds = need_y ? a->bounds_y : nullptr; p.kk = a->kk_y; p.ksize = a->ksize_y;
        p.out_u8 = a->out_u8; p.out = a->out;
        for (int c = 0; c < 3; ++c) { p.mean[c] = a->mean[c]; p.stdv[c] = a->std[c]; }
        dim3 grid((a->out_w + 127) / 128, a->out_h);
        resample_cols_normalize_kernel<<<grid, 128, 0, st>>>(p);
        return check_launch("resample_cols_normalize");
    }
}
