// Depthwise 5x5 + folded BN + ReLU of the ConvRefiner blocks (create_block, romatch/models/matcher.py:92-122) for the
// wide 16-bit maps (C = 569 / 1137 / 1377 at strides 4 / 8 / 16), TMA-fed and persistent.
//
// The kernel is FP32-FMA bound by construction (25 FMA per output element, issued as packed FFMA2 on channel pairs),
// so everything that is not an FMA is kept out of the compute warps:
//   * the 12x20 pixel x 64 channel input window of an 8x16 tile is ONE cp.async.bulk.tensor.4d issued by a loader
//     thread into a 3-stage shared-memory ring (the image border and the channel tail are the tensor map's
//     out-of-bounds zero fill: no address arithmetic, no bounds checks, no staging registers);
//   * CTAs are persistent per 64-channel group, so the 2 x 25 filter taps of a lane's channel pair are loaded once;
//   * 4 compute warps each own two output rows of the tile: 120 LDS.32 + 800 FFMA2 per 64 outputs.
// ncu on the previous version (software loads in the compute warps): 2690 instructions per warp and tile of which 800
// FFMA2, FMA pipe 50 % busy, 115 us for the 216x216x569 map; see DESIGN.md for the numbers of this one.
#include "common.cuh"
#include <cuda.h>

namespace rb {
namespace dwt {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
}  // namespace dwt

constexpr int DT_TH = 8, DT_TW = 16, DT_CH = 64, DT_IH = DT_TH + 4, DT_IW = DT_TW + 4;
constexpr int DT_STAGES = 3;
constexpr int DT_STAGE_BYTES = DT_IH * DT_IW * DT_CH * 2;            // 30720
constexpr int DT_THREADS = 128 + 32;                                  // 4 compute warps + loader warp
constexpr int DT_SMEM = DT_STAGES * DT_STAGE_BYTES + 128 + 128;       // ring + alignment slack + barriers

#ifdef RB_FZ_CLK
__device__ long long g_dw_clk[16];
#endif

struct DwTmaParams {
    void* out; int64_t ldo;
    const float* wgt; int64_t ldw; const float* bias;
    int H, W, C, tiles_x, tiles_per_img, total_tiles;
};

template <typename T>
__global__ void __launch_bounds__(DT_THREADS) dwconv5x5_relu_tma_kernel(const __grid_constant__ CUtensorMap map_in, const DwTmaParams p) {
    using namespace dwt;
    extern __shared__ uint8_t dsm_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + DT_STAGES * DT_STAGE_BYTES);
    uint64_t* empty = full + DT_STAGES;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    rb::pdl_wait();                                    // everything above overlapped the previous kernel's tail
    const int c0 = blockIdx.y * DT_CH;

    if (wid == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int s = it % DT_STAGES, round = it / DT_STAGES;
                const int img = tile / p.tiles_per_img, r = tile - img * p.tiles_per_img;
                const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                mbar_expect_tx(&full[s], DT_STAGE_BYTES);
                tma_load_4d(ring + s * DT_STAGE_BYTES, &map_in, &full[s], c0, tx * DT_TW - 2, ty * DT_TH - 2, img);
            }
        }
        return;
    }
    const int c = c0 + 2 * lane;
    const bool ok0 = c < p.C, ok1 = c + 1 < p.C;
    float2 wv[25];                                        // (channel c, channel c+1) taps: operands of the packed FFMA2
#pragma unroll
    for (int t = 0; t < 25; ++t) wv[t] = make_float2(ok0 ? p.wgt[(int64_t)t * p.ldw + c] : 0.f, ok1 ? p.wgt[(int64_t)t * p.ldw + c + 1] : 0.f);
    const float2 bv = make_float2(ok0 ? p.bias[c] : 0.f, ok1 ? p.bias[c + 1] : 0.f);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int s = it % DT_STAGES, round = it / DT_STAGES;
        const int img = tile / p.tiles_per_img, r = tile - img * p.tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int x0 = tx * DT_TW, y0 = ty * DT_TH;
        const T* tile_s = reinterpret_cast<const T*>(ring + s * DT_STAGE_BYTES);
#ifdef RB_FZ_CLK
        const long long t0 = clock64();
#endif
        mbar_wait(&full[s], round & 1);
#ifdef RB_FZ_CLK
        const long long t1 = clock64();
#endif
        float2 acc[2][DT_TW];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int i = 0; i < DT_TW; ++i) acc[rr][i] = bv;
#pragma unroll
        for (int iy = 0; iy < 6; ++iy) {                      // input rows 2*wid + iy of the window feed output rows 2*wid + {0, 1}
#pragma unroll
            for (int px = 0; px < DT_IW; ++px) {
                T pr[2];
                *reinterpret_cast<uint32_t*>(pr) = *reinterpret_cast<const uint32_t*>(&tile_s[((2 * wid + iy) * DT_IW + px) * DT_CH + 2 * lane]);
                const float2 v = make_float2(to_f(pr[0]), to_f(pr[1]));
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int ky = iy - rr;
                    if (ky >= 0 && ky < 5) {
#pragma unroll
                        for (int kx = 0; kx < 5; ++kx) {
                            const int ox = px - kx;
                            if (ox >= 0 && ox < DT_TW) acc[rr][ox] = __ffma2_rn(wv[ky * 5 + kx], v, acc[rr][ox]);
                        }
                    }
                }
            }
        }
#ifdef RB_FZ_CLK
        const long long t2 = clock64();
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);                // this warp no longer reads the stage
        if (ok0) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int yy = y0 + 2 * wid + rr;
                if (yy < p.H) {
                    T* ob = (T*)p.out + ((int64_t)img * p.H * p.W + (int64_t)yy * p.W) * p.ldo + c;
#pragma unroll
                    for (int i = 0; i < DT_TW; ++i) {
                        if (x0 + i < p.W) {
                            T pair[2] = {from_f<T>(fmaxf(acc[rr][i].x, 0.f)), from_f<T>(ok1 ? fmaxf(acc[rr][i].y, 0.f) : 0.f)};
                            *reinterpret_cast<uint32_t*>(ob + (int64_t)(x0 + i) * p.ldo) = *reinterpret_cast<uint32_t*>(pair);
                        }
                    }
                }
            }
        }
#ifdef RB_FZ_CLK
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && wid == 0) { const long long t3 = clock64(); g_dw_clk[0] += t1 - t0; g_dw_clk[1] += t2 - t1; g_dw_clk[2] += t3 - t2; g_dw_clk[3] += 1; }
#endif
    }
}

typedef CUresult (*EncodeTiledFnDw)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#ifdef RB_FZ_CLK
extern "C" int romab200_debug_dwclk(long long* out, int reset) {
    if (reset) { long long z[16] = {0}; return (int)cudaMemcpyToSymbol(g_dw_clk, z, sizeof(z)); }
    return (int)cudaMemcpyFromSymbol(out, g_dw_clk, sizeof(long long) * 16);
}
#endif

// 16-bit maps only; the caller has checked ldi % 8 == 0, ldo % 2 == 0 and the 16-byte alignment of the input.
int dwconv_tma(const rb_dwconv_args* a, cudaStream_t st) {
    static EncodeTiledFnDw enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        RB_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && ptr,
                   "dwconv: cuTensorMapEncodeTiled not available");
        enc = (EncodeTiledFnDw)ptr;
    }
    CUtensorMap map;             // activation [B, H, W, C] with pitch ldi: box = 12 x 20 pixels x 64 channels, borders zero-filled
    cuuint64_t d4[4] = {(cuuint64_t)a->c, (cuuint64_t)a->w, (cuuint64_t)a->h, (cuuint64_t)a->batch};
    cuuint64_t s4[3] = {(cuuint64_t)a->ldi * 2, (cuuint64_t)a->w * a->ldi * 2, (cuuint64_t)a->h * a->w * a->ldi * 2};
    cuuint32_t b4[4] = {(cuuint32_t)DT_CH, (cuuint32_t)DT_IW, (cuuint32_t)DT_IH, 1};
    cuuint32_t e4[4] = {1, 1, 1, 1};
    CUresult r4 = enc(&map, a->dtype == RB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(a->in),
                      d4, s4, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RB_REQUIRE(r4 == CUDA_SUCCESS, "dwconv: cuTensorMapEncodeTiled failed with %d (c=%d w=%d h=%d ldi=%lld)", (int)r4, a->c, a->w, a->h, (long long)a->ldi);
    DwTmaParams p;
    p.out = a->out; p.ldo = a->ldo; p.wgt = a->weight; p.ldw = a->ldw; p.bias = a->bias;
    p.H = a->h; p.W = a->w; p.C = a->c;
    p.tiles_x = (a->w + DT_TW - 1) / DT_TW;
    p.tiles_per_img = p.tiles_x * ((a->h + DT_TH - 1) / DT_TH);
    const long long total = (long long)p.tiles_per_img * a->batch;
    RB_REQUIRE(total > 0 && total < (1ll << 31), "dwconv: bad tile count");
    p.total_tiles = (int)total;
    const int groups = (a->c + DT_CH - 1) / DT_CH;
    RB_REQUIRE(groups <= 65535, "dwconv: too many channel groups");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_group = 2 * sms / groups;                          // two resident CTAs per SM, never more CTAs than fit at once
    if (per_group < 1) per_group = 1;
    if (per_group > p.total_tiles) per_group = p.total_tiles;
    dim3 grid(per_group, groups);
    if (a->dtype == RB_F16) {
        static bool cfg = false;
        if (!cfg) { RB_REQUIRE(cudaFuncSetAttribute(dwconv5x5_relu_tma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM) == cudaSuccess, "dwconv: smem attribute"); cfg = true; }
        rb::launch_pdl(dwconv5x5_relu_tma_kernel<__half>, grid, dim3(DT_THREADS), DT_SMEM, st, map, p);
    } else {
        static bool cfg = false;
        if (!cfg) { RB_REQUIRE(cudaFuncSetAttribute(dwconv5x5_relu_tma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM) == cudaSuccess, "dwconv: smem attribute"); cfg = true; }
        rb::launch_pdl(dwconv5x5_relu_tma_kernel<__nv_bfloat16>, grid, dim3(DT_THREADS), DT_SMEM, st, map, p);
    }
    return check_launch("dwconv5x5_relu_tma");
}

}  // namespace rb
