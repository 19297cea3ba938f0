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
