// runtime-chain variant without auxiliary tensors: the K_CHAIN specialisation of chain_rows_kernel (see dmxq_rows.cuh)
#include "dmxq_rows.cuh"

namespace dmxq {

cudaError_t launch_rows_g(int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s)
{
    return launch_rows_kind<K_CHAIN>(in_dt, out_dt, flat, p, s);
}

}  // namespace dmxq
