// tcgen05 / TMA GEMM back-end (16-bit operands, fp32 accumulation in TMEM).  Placeholder until the
// kernel lands: refuses loudly instead of falling back.
#include "common.cuh"
namespace rb {
int gemm_tc(const rb_gemm_args* a, cudaStream_t stream) {
    (void)stream;
    set_error("gemm: tcgen05 back-end not built yet (dtype_ab=%d)", a->dtype_ab);
    return 1;
}
}  // namespace rb
