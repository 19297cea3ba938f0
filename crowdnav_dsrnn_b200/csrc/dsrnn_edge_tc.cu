// dsrnn_edge_tc.cu -- tensor-core (tcgen05) edge-GRU stage of the DS-RNN forward.  PLACEHOLDER until the
// tcgen05 kernel lands: create/destroy are no-ops and the forward reports that the precision is unavailable.
#include "dsrnn.cuh"

const char *dsrnn_tc_create(const CnDsrnnWeights *, cudaStream_t, void **state) { *state = nullptr; return nullptr; }
void dsrnn_tc_destroy(void *) {}
const char *dsrnn_tc_edge_forward(void *, const CnDsrnnWeights *, int, int, const CnDsrnnIO *, int, cudaStream_t, int *)
{
    return "tensor-core edge stage not built yet: use precision fp32";
}
