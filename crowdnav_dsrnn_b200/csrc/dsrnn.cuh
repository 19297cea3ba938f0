// dsrnn.cuh -- internal interface between c_abi.cu and the DS-RNN forward kernels (K3).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/crowdnav_b200.h"

// each returns NULL on success or a static error string
const char *dsrnn_create(const CnDsrnnWeights *w, int device, cudaStream_t stream, CnDsrnn **out);
void dsrnn_destroy(CnDsrnn *m);
const char *dsrnn_update_weights(CnDsrnn *m, const CnDsrnnWeights *w, cudaStream_t stream);
size_t dsrnn_workspace_bytes(int n_envs, int human_num);
const char *dsrnn_forward(CnDsrnn *m, int n_envs, int human_num, const CnDsrnnIO *io, int precision,
                          void *workspace, cudaStream_t stream);
int dsrnn_last_launches(const CnDsrnn *m);
int dsrnn_device(const CnDsrnn *m);
const char *dsrnn_edge_sequence_step(CnDsrnn *m, int n_envs, int human_num, const CnEdgeSeqStep *io, cudaStream_t stream);
void dsrnn_enable_timing(CnDsrnn *m, int enable);
void dsrnn_set_refill_env(CnDsrnn *m, CnEnv *env);
void dsrnn_set_edge_event(CnDsrnn *m, void *event);
void dsrnn_set_edge_image(CnDsrnn *m, void *in_hi, void *in_lo, void *out_hi, void *out_lo);
float dsrnn_time_ms(CnDsrnn *m, int *count);

// ---- programmatic dependent launch (PDL) between the kernels of one forward -------------------------------------------
// A kernel launched with cn_launch(..., pdl = true) may become resident while its predecessor in the stream is still
// draining: its prologue (barrier init, TMEM allocation, constant staging, first weight chunks) overlaps the predecessor's
// tail.  Rules kept by every kernel that takes part: (1) pdl_wait() comes before the first access to anything the
// predecessor writes or reads, (2) pdl_launch_dependents() comes AFTER pdl_wait(), so a kernel can only start once the
// kernel two places before it has completed -- what runs before pdl_wait() may read data that is at least two kernels old
// (packed weights), nothing newer.  Without the launch attribute both instructions are no-ops.  The attribute is OFF by
// default: the prologues turned out to be a few microseconds against dependent chains of 16-480 us, and an early-resident
// attention grid delays the spare-episode refill into the node kernel (profiles/r2_pdl_experiment.txt).  CN_PDL=<mask>
// switches it on per kernel (same results; development A/B switch).
enum { CN_PDL_EDGE = 1, CN_PDL_LINEAR = 2, CN_PDL_ATTENTION = 4, CN_PDL_NODE = 8 };
int cn_pdl_mask();          // CN_PDL=<bit mask>; default 0 = off: measured, no gain (profiles/r2_pdl_experiment.txt)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t cn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int pdl_bit, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (cn_pdl_mask() & pdl_bit) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif
