// many-tensor launches of chain_rows_kernel's body (dmxq_cast_chain_multi): the kinds a whole-model weight cast uses
#include "dmxq_rows.cuh"

namespace dmxq {

bool rows_multi_supported(int kind)
{
    return kind == K_BFP || kind == K_SBFP || kind == K_NM_BFP || kind == K_NM24_BFP || kind == K_NM || kind == K_FLOAT || kind == K_CHAIN;
}

template <int KIND> static cudaError_t multi_kind(int dt, const RowsParams &p, const MultiTable &t, cudaStream_t s)
{
    if (dt == 0) return launch_rows_multi_k<float, KIND>(p, t, s);
    if (dt == 1) return launch_rows_multi_k<__nv_bfloat16, KIND>(p, t, s);
    if (dt == 2) return launch_rows_multi_k<__half, KIND>(p, t, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_rows_multi(int dt, int kind, const RowsParams &p, const MultiTable &t, cudaStream_t s)
{
    switch (kind) {
    case K_BFP: return multi_kind<K_BFP>(dt, p, t, s);
    case K_SBFP: return multi_kind<K_SBFP>(dt, p, t, s);
    case K_NM_BFP: return multi_kind<K_NM_BFP>(dt, p, t, s);
    case K_NM: return multi_kind<K_NM>(dt, p, t, s);
    case K_FLOAT: return multi_kind<K_FLOAT>(dt, p, t, s);
    case K_CHAIN: return multi_kind<K_CHAIN>(dt, p, t, s);
    case K_NM24_BFP:
        if (dt == 1) return launch_rows_multi_k<__nv_bfloat16, K_NM24_BFP>(p, t, s);
        if (dt == 2) return launch_rows_multi_k<__half, K_NM24_BFP>(p, t, s);
        return cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace dmxq
