// runtime-chain variants of chain_rows_kernel (see dmxq_rows.cuh) + the KIND dispatcher
#include "dmxq_rows.cuh"

namespace dmxq {

cudaError_t launch_rows_b(int kind, int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s);
cudaError_t launch_rows_c(int kind, int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s);
cudaError_t launch_rows_d(int kind, int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s);
cudaError_t launch_rows_e(int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s);
cudaError_t launch_rows_g(int in_dt, int out_dt, bool flat, const RowsParams &p, cudaStream_t s);

cudaError_t launch_rows(int in_dt, int out_dt, bool flat, int kind, const RowsParams &p, cudaStream_t s)
{
    switch (kind) {
    case K_AUX: return launch_rows_kind<K_AUX>(in_dt, out_dt, flat, p, s);
    case K_CHAIN: return launch_rows_g(in_dt, out_dt, flat, p, s);
    case K_BFP:
    case K_FLOAT: return launch_rows_b(kind, in_dt, out_dt, flat, p, s);
    case K_MXFP:
    case K_BFP_ASYM:
    case K_BFP_STOCH: return launch_rows_d(kind, in_dt, out_dt, flat, p, s);
    case K_NM24_BFP: return launch_rows_e(in_dt, out_dt, flat, p, s);
    default: return launch_rows_c(kind, in_dt, out_dt, flat, p, s);
    }
}

}  // namespace dmxq
