// Pieces shared by the refiner prologue kernels (refiner.cu: one warp per pixel; local_corr_tile.cu: one CTA per
// tile of pixels whose windows overlap).  romatch/models/matcher.py:132-168, romatch/utils/local_correlation.py:77-142.
#pragma once
#include "common.cuh"

namespace rb {

struct PrologueParams {
    const void* feat; int64_t ldf; int n_img, y_shift;
    const float* state; void* d; int64_t ldd;
    int D, h, w, cf, emb;
    const float* emb_w; const float* emb_b; float disp_scale;
    const float* gx; const float* gy; const float* winx; const float* winy;
    int vec_ok;
    const unsigned char* tile_done;     // per tile of LcTile<R>: 1 = the tile kernel wrote these pixels (nullptr: no tile pass ran)
    const float* corr_table; int64_t ld_table;   // optional all-pairs table [D*h*w][ld_table]: scale * <x[item, p], y[item, q]> for every position q
};

// tile geometry of the cooperative prologue (local_corr_tile.cu) per window radius: TQX x TQY query pixels per CTA, one thread
// per (pixel, window row), channels staged CK at a time, at most MAXPOS positions of f1 (the union of the tile's windows,
// rows padded to a multiple of 8 positions) per stage
template <int R> struct LcTile { static constexpr int TQX = 8, TQY = 4, CK = 16, MAXPOS = 384; };
template <> struct LcTile<7> { static constexpr int TQX = 8, TQY = 2, CK = 16, MAXPOS = 672; };
template <> struct LcTile<2> { static constexpr int TQX = 8, TQY = 4, CK = 16, MAXPOS = 336; };
template <> struct LcTile<0> { static constexpr int TQX = 8, TQY = 4, CK = 16, MAXPOS = 8; };

template <int R>
__host__ __device__ __forceinline__ int lc_tile_index(int item, int y, int x, int h, int w) {
    const int tx = (w + LcTile<R>::TQX - 1) / LcTile<R>::TQX, ty = (h + LcTile<R>::TQY - 1) / LcTile<R>::TQY;
    return (item * ty + y / LcTile<R>::TQY) * tx + x / LcTile<R>::TQX;
}

// The (2R+1)^2 window samples of one pixel from its table D[j][i] = scale * <f0, f1[by+j, bx+i]> on the (2R+2)^2 integer
// neighbourhood: every sample is the bilinear blend of four D entries with the weights grid_sample would use for the
// coordinate flow + window[k] (all samples sit on a unit pixel lattice).  One warp; lanes over k.
template <int R, typename TO>
__device__ __forceinline__ void lc_blend_window(const float* __restrict__ dtab, float fx, float fy, int bx, int by, int h, int w,
                                                const float* __restrict__ winx, const float* __restrict__ winy, TO* __restrict__ out, int lane) {
    constexpr int S = 2 * R + 2, K1 = 2 * R + 1, K = K1 * K1;
    for (int k = lane; k < K; k += 32) {
        int dy = k / K1, dx = k - dy * K1;
        float xk = fx + winx[dx], yk = fy + winy[dy];
        float ix = ((xk + 1.f) * w - 1.f) * 0.5f, iy = ((yk + 1.f) * h - 1.f) * 0.5f;
        float x0 = floorf(ix), y0 = floorf(iy);
        float wx1 = ix - x0, wx0 = (x0 + 1.f) - ix, wy1 = iy - y0, wy0 = (y0 + 1.f) - iy;
        int ti = (int)x0 - bx, tj = (int)y0 - by;
        int ti0 = min(max(ti, 0), S - 1), ti1 = min(max(ti + 1, 0), S - 1);
        int tj0 = min(max(tj, 0), S - 1), tj1 = min(max(tj + 1, 0), S - 1);
        float v = dtab[tj0 * S + ti0] * (wx0 * wy0) + dtab[tj0 * S + ti1] * (wx1 * wy0) +
                  dtab[tj1 * S + ti0] * (wx0 * wy1) + dtab[tj1 * S + ti1] * (wx1 * wy1);
        out[k] = from_f<TO>(v);
    }
}

int refiner_prologue_tile(const PrologueParams& p, int radius, unsigned char* tile_done, cudaStream_t st);   // local_corr_tile.cu

}  // namespace rb
