// MXFP specialisation of chain_rows_kernel: K_MXFP (see dmxq_rows.cuh)
#include "dmxq_rows.cuh"

namespace dmxq {

cudaError_t launch_rows_d(int kind, int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s)
{
    (void)kind;
    return launch_rows_kind<K_MXFP>(in_dt, out_dt, flat, p, s);
}

}  // namespace dmxq
