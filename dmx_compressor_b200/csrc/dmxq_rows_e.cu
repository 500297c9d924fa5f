// 2:4 -> BFP on 16-bit tensors: the K_NM24_BFP specialisation of chain_rows_kernel (see dmxq_rows.cuh)
#include "dmxq_rows.cuh"

namespace dmxq {

cudaError_t launch_rows_e(int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s)
{
    if (in_dt == 1 && out_dt == 1) return launch_rows_k<__nv_bfloat16, __nv_bfloat16, K_NM24_BFP>(flat, p, s);
    if (in_dt == 2 && out_dt == 2) return launch_rows_k<__half, __half, K_NM24_BFP>(flat, p, s);
    return cudaErrorInvalidValue;  // (the dispatcher only sends same-dtype 16-bit tensors here)
}

}  // namespace dmxq
