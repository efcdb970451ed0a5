// Weighted sampling without replacement on the device — the two `torch.multinomial(..., replacement=False)` draws of
// `RegressionMatcher.sample` (romatch/models/matcher.py:613-617, 626-628), SURVEY 8f-1.
//
// Exponential race (Efraimidis-Spirakis): key_i = -log(u_i) / w_i with u_i ~ U(0,1); the k smallest keys are a draw of k items
// without replacement with probabilities proportional to w (the same construction torch.multinomial uses on CUDA: w / Exp(1),
// top-k).  One CTA per batch item:
//   pass 0  keys from a counter-based generator (Philox4x32-10 keyed by the seed, counter = element index), written to the
//           workspace; the weight transform of the caller is applied on the fly (certainty threshold / density balancing), so
//           no intermediate tensor is materialised;
//   pass 1-3  radix select of the k-th smallest key: shared-memory histograms over bits [31:21], [20:10], [9:0] of the
//           (order-preserving) bit pattern of the positive float keys;
//   pass 4  compaction: indices of all keys below the threshold and as many ties as are still needed.
// The keys of a 1.5 M-pixel certainty map are 6 MB: L2-resident across the passes.  Items of zero weight have key = +inf and are
// only drawn when fewer than k positive weights exist.
#include "common.cuh"

namespace rb {

__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
    const uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0, c1, 0, 0), key (k0, k1); returns the first output word
__device__ __forceinline__ uint32_t philox_u32(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
    uint32_t c[4] = {c0, c1, 0u, 0u};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, hi1;
        const uint32_t lo0 = mulhilo32(0xD2511F53u, c[0], &hi0);
        const uint32_t lo1 = mulhilo32(0xCD9E8D57u, c[2], &hi1);
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c[0];
}

// weight transforms (matcher.py:604-607, 622-625)
__device__ __forceinline__ float sample_weight(float v, int transform, float param) {
    if (transform == RB_SAMPLE_THRESHOLD) return v > param ? 1.0f : v;                       // certainty[certainty > thresh] = 1
    if (transform == RB_SAMPLE_BALANCE) return v < 10.0f ? 1e-7f : 1.0f / (v + 1.0f);        // p = 1/(density+1); p[density < 10] = 1e-7
    return v;
}

constexpr int SMP_THREADS = 1024;

__global__ void __launch_bounds__(SMP_THREADS) weighted_sample_kernel(const float* __restrict__ values, int64_t n, int k, int64_t stride, uint64_t seed,
                                                                      int transform, float param, int32_t* __restrict__ out_idx, float* __restrict__ out_w,
                                                                      float* __restrict__ keys_ws) {
    rb::pdl_wait();
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t s_prefix, s_need, s_count, s_ties;
    const int tid = threadIdx.x, b = blockIdx.x;
    const float* v = values + (int64_t)b * stride;
    float* keys = keys_ws + (int64_t)b * n;
    int32_t* out = out_idx + (int64_t)b * k;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)b * 0x9E3779B9u;
    // pass 0: keys
    for (int64_t i = tid; i < n; i += SMP_THREADS) {
        const float w = sample_weight(v[i], transform, param);
        const uint32_t r = philox_u32((uint32_t)i, (uint32_t)(i >> 32), k0, k1);
        const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);          // (0, 1), 24 random bits
        keys[i] = w > 0.f ? -__logf(u) / w : __int_as_float(0x7f800000);
    }
    __syncthreads();
    // passes 1-3: radix select on the bit patterns (positive floats and +inf order like unsigned integers)
    uint32_t prefix = 0, need = (uint32_t)k;         // keys whose high bits equal `prefix` are candidates; `need` = rank inside them
    const int shifts[3] = {21, 10, 0}, nbits[3] = {11, 11, 10};
    uint32_t mask = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = shifts[pass], bins = 1 << nbits[pass];
        for (int i = tid; i < bins; i += SMP_THREADS) hist[i] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n; i += SMP_THREADS) {
            const uint32_t x = __float_as_uint(keys[i]);
            if ((x & mask) == prefix) atomicAdd(&hist[(x >> sh) & (bins - 1)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t cum = 0; int sel = bins - 1;
            for (int i = 0; i < bins; ++i) {
                if (cum + hist[i] >= need) { sel = i; break; }
                cum += hist[i];
            }
            s_prefix = prefix | ((uint32_t)sel << sh);
            s_need = need - cum;
        }
        __syncthreads();
        prefix = s_prefix; need = s_need;
        mask |= (uint32_t)(bins - 1) << sh;
        __syncthreads();
    }
    // prefix = bit pattern of the k-th smallest key; `need` = how many keys equal to it are still to be taken
    if (tid == 0) { s_count = 0; s_ties = 0; }
    __syncthreads();
    for (int64_t i = tid; i < n; i += SMP_THREADS) {
        const uint32_t x = __float_as_uint(keys[i]);
        bool take = x < prefix;
        if (x == prefix) take = atomicAdd(&s_ties, 1u) < need;
        if (take) {
            const uint32_t pos = atomicAdd(&s_count, 1u);
            if (pos < (uint32_t)k) {
                out[pos] = (int32_t)i;
                if (out_w) out_w[(int64_t)b * k + pos] = sample_weight(v[i], transform, param);
            }
        }
    }
}

}  // namespace rb

using namespace rb;

extern "C" int romab200_weighted_sample(const rb_sample_args* a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    RB_REQUIRE(a->values && a->out_idx && a->keys, "weighted_sample: null argument");
    RB_REQUIRE(a->n > 0 && a->k > 0 && a->k <= a->n && a->n < (1ll << 31) && a->batch > 0 && a->batch <= 65535, "weighted_sample: bad shape n=%lld k=%d batch=%d",
               (long long)a->n, a->k, a->batch);
    RB_REQUIRE(a->transform >= RB_SAMPLE_IDENTITY && a->transform <= RB_SAMPLE_BALANCE, "weighted_sample: unknown transform %d", a->transform);
    rb::launch_pdl(weighted_sample_kernel, dim3(a->batch), dim3(SMP_THREADS), 0, st, a->values, a->n, a->k, a->stride > 0 ? a->stride : a->n, a->seed, a->transform,
                   a->param, a->out_idx, a->out_weights, a->keys);
    return check_launch("weighted_sample");
}
