// dsrnn_tc_linear.cuh -- host interface of the tcgen05 tall-skinny linear layer (dsrnn_tc_linear.cu).
#pragma once
#include <cuda_runtime.h>

struct TcLinear {
    void *wimg = nullptr;     // packed split-bf16 weight images
    float *bias = nullptr;    // zero-padded bias [n_blocks * n_tile]
    int N = 0, K = 0, n_tile = 0, n_blocks = 0;
};

struct TcLinearCall {
    const float *X; int ldx;
    int rows_per_env = 1, env_stride_rows = 1, first_row = 0;   // memory row of logical row m
    const float *rowscale = nullptr;                            // optional per-env multiplier of the input rows
    int M;
    float *Y; int ldy; int ycol0 = 0;
    int act = 0;                                                // 0 none, 1 ReLU, 2 tanh
    int three_pass = 1;
    int fp16 = 0;                                               // single pass with FP16 operands
};

// W = [W0 (n0 rows); W1 (n1 rows, may be NULL/0)], row-major [*, K]; each returns NULL or an error string
const char *tc_linear_create(TcLinear *L, const float *W0, const float *b0, int n0, const float *W1, const float *b1, int n1,
                             int K, int n_tile, cudaStream_t stream);
const char *tc_linear_repack(TcLinear *L, const float *W0, const float *b0, int n0, const float *W1, const float *b1, int n1,
                             cudaStream_t stream);
void tc_linear_destroy(TcLinear *L);
const char *tc_linear_run(const TcLinear *L, const TcLinearCall &c, int num_sms, cudaStream_t stream);
