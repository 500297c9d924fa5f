// MXFP and asymmetric-BFP specialisations of chain_rows_kernel: K_MXFP, K_BFP_ASYM, K_BFP_STOCH (see dmxq_rows.cuh)
#include "dmxq_rows.cuh"

namespace dmxq {

cudaError_t launch_rows_d(int kind, int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s)
{
    if (kind == K_BFP_ASYM) return launch_rows_kind<K_BFP_ASYM>(in_dt, out_dt, flat, p, s);
    if (kind == K_BFP_STOCH) return launch_rows_kind<K_BFP_STOCH>(in_dt, out_dt, flat, p, s);
    return launch_rows_kind<K_MXFP>(in_dt, out_dt, flat, p, s);
}

}  // namespace dmxq
