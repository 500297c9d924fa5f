// fused / scaled specialisations of chain_rows_kernel: K_FLOAT_BFP, K_NM_BFP, K_SBFP (see dmxq_rows.cuh)
#include "dmxq_rows.cuh"

namespace dmxq {

cudaError_t launch_rows_c(int kind, int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s)
{
    if (kind == K_FLOAT_BFP) return launch_rows_kind<K_FLOAT_BFP>(in_dt, out_dt, flat, p, s);
    if (kind == K_NM_BFP) return launch_rows_kind<K_NM_BFP>(in_dt, out_dt, flat, p, s);
    if (kind == K_FIXED) return launch_rows_kind<K_FIXED>(in_dt, out_dt, flat, p, s);
    if (kind == K_NM) return launch_rows_kind<K_NM>(in_dt, out_dt, flat, p, s);
    return launch_rows_kind<K_SBFP>(in_dt, out_dt, flat, p, s);
}

}  // namespace dmxq
