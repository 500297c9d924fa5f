// single-format specialisations of chain_rows_kernel: K_BFP, K_FLOAT (see dmxq_rows.cuh)
#include "dmxq_rows.cuh"

namespace dmxq {

cudaError_t launch_rows_b(int kind, int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s)
{
    if (kind == K_BFP) return launch_rows_kind<K_BFP>(in_dt, out_dt, flat, p, s);
    return launch_rows_kind<K_FLOAT>(in_dt, out_dt, flat, p, s);
}

}  // namespace dmxq
