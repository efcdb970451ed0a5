// GP posterior solve (romatch/models/matcher.py:301-309): batched fp32 Cholesky of K_yy + sigma*I
// and the two triangular solves with the Fourier basis as right-hand sides, without leaving the GPU
// and without cuSOLVER.
//
// Blocked right-looking factorisation with 32-wide panels on an augmented workspace
//      W = [ K_yy + sigma I ]   n rows
//          [      F^T       ]   nrhs rows
// Applying the panel solve and trailing update to the F^T rows as well turns them into (L^-1 F)^T, so the
// forward substitution is free.  The backward substitution then runs row-wise on those rows
// (X^T L = Y^T), again panel by panel, and leaves alpha^T = X^T in place: exactly the [N,K] operand
// layout that mu = K_xy @ alpha needs.  Panel kernels are latency-bound (one thread per row, 32-step
// recurrences held in registers); all O(n^3) work is in the trailing updates, which are romab200 GEMMs.
#include "common.cuh"

namespace rb {

constexpr int NB = 32;

// one warp factors the (bs x bs, bs <= 32) diagonal block held one row per lane; result (lower) to smem L
__device__ void factor_diag_block(const float* __restrict__ Wd, int64_t ldw, int bs, float (*L)[NB + 1], float* Wout) {
    const int lane = threadIdx.x & 31;
    float r[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        float v = 0.f;
        if (lane < bs && c < bs && c <= lane) v = Wd[(int64_t)lane * ldw + c];
        if (c == lane && lane >= bs) v = 1.f;      // identity padding
        r[c] = v;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        float djj = __shfl_sync(0xffffffffu, r[j], j);
        float d = sqrtf(djj);
        float lij = 0.f;
        if (lane == j) { r[j] = d; }
        else if (lane > j) { lij = r[j] / d; r[j] = lij; }
#pragma unroll
        for (int c = j + 1; c < NB; ++c) {
            float lcj = __shfl_sync(0xffffffffu, lij, c);
            if (c <= lane) r[c] = fmaf(-lij, lcj, r[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        L[lane][c] = (c <= lane) ? r[c] : 0.f;
        if (Wout && lane < bs && c < bs && c <= lane) Wout[(int64_t)lane * ldw + c] = r[c];
    }
}

// one warp per problem: factor the diagonal block in place
__global__ void __launch_bounds__(32) chol_diag_kernel(float* __restrict__ W, int64_t ldw, int64_t stride, int k, int bs) {
    __shared__ float L[NB][NB + 1];
    float* Wd = W + (int64_t)blockIdx.x * stride + (int64_t)k * ldw + k;
    factor_diag_block(Wd, ldw, bs, L, Wd);
}

__device__ __forceinline__ void load_diag_block(const float* __restrict__ Wd, int64_t ldw, int bs, float (*L)[NB + 1]) {
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        int i = idx / NB, c = idx % NB;
        float v = 0.f;
        if (i < bs && c < bs && c <= i) v = Wd[(int64_t)i * ldw + c];
        if (i == c && i >= bs) v = 1.f;        // identity padding of a partial block
        L[i][c] = v;
    }
}

// grid: (row_blocks, batch); block 128 threads: one thread per row below the (already factored) diagonal
// block solves x * L^T = a with the 32-step recurrence held in registers.
__global__ void __launch_bounds__(128) chol_panel_kernel(float* __restrict__ W, int64_t ldw, int64_t stride, int k, int bs, int total_rows) {
    __shared__ float L[NB][NB + 1];
    float* Wb = W + (int64_t)blockIdx.y * stride;
    load_diag_block(Wb + (int64_t)k * ldw + k, ldw, bs, L);
    __syncthreads();
    int row = k + bs + blockIdx.x * 128 + threadIdx.x;
    if (row >= total_rows) return;
    float* ar = Wb + (int64_t)row * ldw + k;
    float a[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) a[c] = c < bs ? ar[c] : 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        float x = a[j] / L[j][j];
        a[j] = x;
#pragma unroll
        for (int c = j + 1; c < NB; ++c) a[c] = fmaf(-x, L[c][j], a[c]);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c)
        if (c < bs) ar[c] = a[c];
}

// backward substitution panel: rows of the RHS block solve x * L_kk = y (L_kk lower, bs x bs)
__global__ void __launch_bounds__(128) trsm_back_kernel(float* __restrict__ W, int64_t ldw, int64_t stride, int k, int bs, int n, int nrhs) {
    __shared__ float L[NB][NB + 1];
    float* Wb = W + (int64_t)blockIdx.y * stride;
    load_diag_block(Wb + (int64_t)k * ldw + k, ldw, bs, L);
    __syncthreads();
    int r = blockIdx.x * 128 + threadIdx.x;
    if (r >= nrhs) return;
    float* yr = Wb + (int64_t)(n + r) * ldw + k;
    float y[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) y[c] = c < bs ? yr[c] : 0.f;
#pragma unroll
    for (int j = NB - 1; j >= 0; --j) {
        float x = y[j] / L[j][j];
        y[j] = x;
#pragma unroll
        for (int c = 0; c < j; ++c) y[c] = fmaf(-x, L[j][c], y[c]);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c)
        if (c < bs) yr[c] = y[c];
}

static int sub_gemm(const float* A, int64_t lda, const float* B, int64_t ldb, int trans_b, float* C, int64_t ldc, int M, int N, int K,
                    int batch, int64_t stride, cudaStream_t st) {
    rb_gemm_args g = {};
    g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.dtype_ab = RB_F32; g.dtype_c = RB_F32; g.trans_b = trans_b;
    g.batch0 = batch; g.batch1 = 1; g.sa0 = stride; g.sb0 = stride; g.sc0 = stride; g.sr0 = stride;
    g.ntaps = 1; g.alpha = -1.0f; g.R = C; g.ldr = ldc; g.dtype_r = RB_F32;
    return gemm_simt(&g, st);
}

}  // namespace rb

using namespace rb;

extern "C" int romab200_gp_solve(const rb_gp_solve_args* a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    RB_REQUIRE(a->n > 0 && a->nrhs > 0 && a->batch > 0 && a->ldw >= a->n, "gp_solve: bad shape n=%d nrhs=%d batch=%d", a->n, a->nrhs, a->batch);
    RB_REQUIRE(a->batch <= 65535, "gp_solve: batch too large");
    const int n = a->n, total = a->n + a->nrhs;
    float* W = a->W;
    // factorisation + forward substitution on the augmented rows
    for (int k = 0; k < n; k += NB) {
        int bs = n - k < NB ? n - k : NB;
        int below = total - (k + bs);
        chol_diag_kernel<<<a->batch, 32, 0, st>>>(W, a->ldw, a->stride, k, bs);
        if (check_launch("chol_diag")) return 1;
        dim3 grid((below + 127) / 128, a->batch);
        chol_panel_kernel<<<grid, 128, 0, st>>>(W, a->ldw, a->stride, k, bs, total);
        if (check_launch("chol_panel")) return 1;
        int nt = n - (k + bs);
        if (nt > 0) {
            float* P = W + (int64_t)(k + bs) * a->ldw + k;           // panel rows below the block
            float* T = W + (int64_t)(k + bs) * a->ldw + (k + bs);    // trailing matrix
            if (sub_gemm(P, a->ldw, P, a->ldw, 0, T, a->ldw, total - (k + bs), nt, bs, a->batch, a->stride, st)) return 1;
        }
    }
    // backward substitution: X^T L = Y^T on rows n .. n+nrhs
    int last = ((n - 1) / NB) * NB;
    for (int k = last; k >= 0; k -= NB) {
        int bs = n - k < NB ? n - k : NB;
        dim3 grid((a->nrhs + 127) / 128, a->batch);
        trsm_back_kernel<<<grid, 128, 0, st>>>(W, a->ldw, a->stride, k, bs, n, a->nrhs);
        if (check_launch("trsm_back")) return 1;
        if (k > 0) {
            float* X = W + (int64_t)n * a->ldw + k;                  // solved block  [nrhs, bs]
            float* Lr = W + (int64_t)k * a->ldw;                     // L[k:k+bs, 0:k] as [K, N]
            float* Y = W + (int64_t)n * a->ldw;                      // remaining RHS [nrhs, k]
            if (sub_gemm(X, a->ldw, Lr, a->ldw, 1, Y, a->ldw, a->nrhs, k, bs, a->batch, a->stride, st)) return 1;
        }
    }
    return 0;
}
