// dsrnn_forward.cu -- K3: DS-RNN policy forward for one rollout step (sm_100a).
//
// SRNN.forward(infer=True) + DiagGaussian.fc_mean
// (pytorchBaselines/a2c_ppo_acktr/srnn_model.py:409-504, distributions.py:85-94):
//   stage 1  edge GRUs: rows = N temporal edges + N*H spatial edges, per row
//            e = ReLU(W_enc x + b), GRU(64 -> 256) on the masked hidden state        (srnn_model.py:201-215, 37-50)
//   stage 2  attention: q = W_t o_t, k_i = W_s o_i, softmax_i(q.k_i * H/sqrt(64)), c = sum a_i o_i   (:256-339)
//   stage 3  node: x = [ReLU(W_ne (W_r robot + b) + b) | ReLU(W_na [o_t|c] + b)], GRU(128 -> 128), y = W_out h  (:149-173)
//   stage 4  heads: actor / critic tanh MLPs, critic_linear, fc_mean                 (:378-395, 487-495)
//
// This file holds the fp32 CUDA-core implementation (precision CN_PREC_FP32: the exact mode and the
// in-library reference for the tensor-core edge stage in dsrnn_edge_tc.cu) and the stage 2-4 kernels
// every precision shares.
#include <cstdio>
#include <new>
#include "dsrnn.cuh"
#include "dsrnn_tc_linear.cuh"

#include <cstdlib>
#include <vector>

// dsrnn_node_tc.cu: stages 3 + 4 as one tcgen05 kernel
const char *dsrnn_node_tc_create(const CnDsrnnWeights *w, cudaStream_t stream, void **state);
const char *dsrnn_node_tc_repack(void *state, const CnDsrnnWeights *w, cudaStream_t stream);
void dsrnn_node_tc_destroy(void *state);
const char *dsrnn_node_tc_forward(void *state, int n_envs, const CnDsrnnIO *io, const float *cat, float *feat, int precision,
                                  cudaStream_t stream, int *launches);

struct CnDsrnn {
    CnDsrnnWeights w;
    int device;
    int last_launches;
    void *tc_state;   // packed bf16 weights of the tensor-core edge stage (dsrnn_edge_tc.cu)
    void *node_state = nullptr;   // packed weights of the fused node / heads kernel (dsrnn_node_tc.cu)
    bool unfused_node = false;    // CN_NODE_UNFUSED=1: one launch per layer (development A/B switch)
    CnEnv *refill_env = nullptr;  // cn_dsrnn_set_refill_env: start this env's spare-episode refill beside the attention kernel
    cudaEvent_t edge_done = nullptr;   // cn_dsrnn_set_edge_event: recorded behind the edge stage of the next forwards
    void *edge_img[4] = {nullptr, nullptr, nullptr, nullptr};   // cn_dsrnn_set_edge_image: resident split-bf16 image of the edge state
    bool timing;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending, pool;   // events around the edge stage
    int num_sms;
    // tensor-core images of the stage 2-4 linears (dsrnn_tc_linear.cu)
    TcLinear att_qt, emb, gi, gh, out, ac0, actor2, critic2;
    float *att_wc = nullptr, *att_bc = nullptr;   // folded attention projection (W_s^T W_t, W_s^T b_t)
};

void dsrnn_enable_timing(CnDsrnn *m, int enable) { dsrnn_time_ms(m, nullptr); m->timing = enable != 0; }

float dsrnn_time_ms(CnDsrnn *m, int *count)
{
    float total = 0.f;
    for (auto &p : m->pending) {
        cudaEventSynchronize(p.second);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p.first, p.second);
        total += ms;
        m->pool.push_back(p);
    }
    if (count) *count = (int)m->pending.size();
    m->pending.clear();
    return total;
}

// implemented in dsrnn_edge_tc.cu
const char *dsrnn_tc_create(const CnDsrnnWeights *w, cudaStream_t stream, void **state);
const char *dsrnn_tc_repack(void *state, const CnDsrnnWeights *w, cudaStream_t stream);
void dsrnn_tc_destroy(void *state);
const char *dsrnn_tc_edge_forward(void *state, const CnDsrnnWeights *w, int n_envs, int H, const CnDsrnnIO *io,
                                  int precision, cudaStream_t stream, int *launches, void *const img[4]);

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

int cn_pdl_mask()
{
    static const int mask = [] { const char *e = getenv("CN_PDL"); return e ? atoi(e) : 0; }();
    return mask;
}

// ---------------------------------------------------------------------------------------------- generic linear
struct LinArgs {
    const float *X; int ldx;
    int rows_per_env, env_stride_rows, first_row;   // memory row of logical row m: (m / rpe) * stride + first + m % rpe
    const float *rowscale;                          // optional [M / rows_per_env] multiplier of the input rows (the mask)
    const float *W; const float *b;                 // W [N, K] row-major
    int M, N, K;
    float *Y; int ldy; int ycol0;
    int act;
};

#define LBM 64
#define LBN 64
#define LBK 16

// Y[m, ycol0 + n] = act(sum_k X[row(m), k] * W[n, k] + b[n]); K % 16 == 0, 16-byte aligned rows
__global__ void __launch_bounds__(256) linear_simt_kernel(const LinArgs a)
{
    __shared__ float As[LBK][LBM + 4];
    __shared__ float Bs[LBK][LBN + 4];
    const int t = threadIdx.x;
    const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * LBN;
    const int lr = t >> 2, lk = (t & 3) * 4;
    const int tx = t & 15, ty = t >> 4;
    float acc[4][4] = {};
    const int am = m0 + lr;
    const bool a_ok = am < a.M;
    size_t a_row = 0;
    float a_scale = 1.0f;
    if (a_ok) {
        const int env = am / a.rows_per_env;
        a_row = (size_t)env * a.env_stride_rows + a.first_row + (am - env * a.rows_per_env);
        if (a.rowscale) a_scale = a.rowscale[env];
    }
    const int bn = n0 + lr;
    const bool b_ok = bn < a.N;
    for (int k0 = 0; k0 < a.K; k0 += LBK) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (a_ok) av = *reinterpret_cast<const float4 *>(a.X + a_row * a.ldx + k0 + lk);
        if (b_ok) bv = *reinterpret_cast<const float4 *>(a.W + (size_t)bn * a.K + k0 + lk);
        As[lk + 0][lr] = av.x * a_scale; As[lk + 1][lr] = av.y * a_scale;
        As[lk + 2][lr] = av.z * a_scale; As[lk + 3][lr] = av.w * a_scale;
        Bs[lk + 0][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LBK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            float v = acc[i][j] + (a.b ? a.b[n] : 0.0f);
            if (a.act == ACT_RELU) v = fmaxf(v, 0.0f);
            else if (a.act == ACT_TANH) v = tanhf(v);
            a.Y[(size_t)m * a.ldy + a.ycol0 + n] = v;
        }
    }
}

static void launch_linear(const LinArgs &a, cudaStream_t s, int *launches)
{
    dim3 grid((a.M + LBM - 1) / LBM, (a.N + LBN - 1) / LBN);
    linear_simt_kernel<<<grid, 256, 0, s>>>(a);
    ++*launches;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------- stage 1 (fp32)
// Edge GRU for one weight set.  Logical row m -> (env = m / rpe, memory row env*(H+1) + first + m % rpe).
struct EdgeArgs {
    const float *x_in;      // [rows, 2] edge features (temporal_edges or spatial_edges, contiguous)
    const float *h_in;      // [N, H+1, 256]
    const float *masks;     // [N]
    const float *enc_w, *enc_b, *w_ih, *w_hh, *b_ih, *b_hh;
    float *h_out;           // [N, H+1, 256]
    int M, rpe, stride, first;
};

// tile: 64 rows x 64 hidden units (x3 gates), K = 64 (encoded input) + 256 (hidden)
__global__ void __launch_bounds__(256) edge_gru_simt_kernel(const EdgeArgs a)
{
    __shared__ float As[LBK][LBM + 4];
    __shared__ float Ws[3][LBK][LBN + 4];
    const int t = threadIdx.x;
    const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * LBN;
    const int lr = t >> 2, lk = (t & 3) * 4;
    const int tx = t & 15, ty = t >> 4;
    float acc_r[4][4] = {}, acc_z[4][4] = {}, acc_ni[4][4] = {}, acc_nh[4][4] = {};
    const int am = m0 + lr;
    const bool a_ok = am < a.M;
    size_t a_row = 0;
    float mask = 0.0f, x0 = 0.0f, x1 = 0.0f;
    if (a_ok) {
        const int env = am / a.rpe;
        a_row = (size_t)env * a.stride + a.first + (am - env * a.rpe);
        mask = a.masks[env];
        x0 = a.x_in[2 * (size_t)am]; x1 = a.x_in[2 * (size_t)am + 1];
    }
    for (int k0 = 0; k0 < 320; k0 += LBK) {
        const bool xr = k0 < 64;
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        if (a_ok) {
            if (xr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = k0 + lk + i;
                    av[i] = fmaxf(fmaf(a.enc_w[2 * k + 1], x1, fmaf(a.enc_w[2 * k], x0, 0.0f)) + a.enc_b[k], 0.0f);
                }
            } else {
                const float4 h4 = *reinterpret_cast<const float4 *>(a.h_in + a_row * 256 + (k0 - 64) + lk);
                av[0] = h4.x * mask; av[1] = h4.y * mask; av[2] = h4.z * mask; av[3] = h4.w * mask;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) As[lk + i][lr] = av[i];
#pragma unroll
        for (int gate = 0; gate < 3; ++gate) {
            const int wrow = gate * 256 + n0 + lr;
            const float4 w4 = xr ? *reinterpret_cast<const float4 *>(a.w_ih + (size_t)wrow * 64 + k0 + lk)
                                 : *reinterpret_cast<const float4 *>(a.w_hh + (size_t)wrow * 256 + (k0 - 64) + lk);
            Ws[gate][lk + 0][lr] = w4.x; Ws[gate][lk + 1][lr] = w4.y; Ws[gate][lk + 2][lr] = w4.z; Ws[gate][lk + 3][lr] = w4.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LBK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 r4 = *reinterpret_cast<const float4 *>(&Ws[0][k][tx * 4]);
            const float4 z4 = *reinterpret_cast<const float4 *>(&Ws[1][k][tx * 4]);
            const float4 n4 = *reinterpret_cast<const float4 *>(&Ws[2][k][tx * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
            const float wr[4] = {r4.x, r4.y, r4.z, r4.w}, wz[4] = {z4.x, z4.y, z4.z, z4.w}, wn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc_r[i][j] = fmaf(ar[i], wr[j], acc_r[i][j]);
                    acc_z[i][j] = fmaf(ar[i], wz[j], acc_z[i][j]);
                    if (xr) acc_ni[i][j] = fmaf(ar[i], wn[j], acc_ni[i][j]);
                    else acc_nh[i][j] = fmaf(ar[i], wn[j], acc_nh[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
        const int env = m / a.rpe;
        const size_t row = (size_t)env * a.stride + a.first + (m - env * a.rpe);
        const float mk = a.masks[env];
        const int c0 = n0 + tx * 4;
        const float4 hp4 = *reinterpret_cast<const float4 *>(a.h_in + row * 256 + c0);
        const float hp[4] = {hp4.x * mk, hp4.y * mk, hp4.z * mk, hp4.w * mk};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            const float r = sigmoidf_(acc_r[i][j] + a.b_ih[c] + a.b_hh[c]);
            const float z = sigmoidf_(acc_z[i][j] + a.b_ih[256 + c] + a.b_hh[256 + c]);
            const float n = tanhf(acc_ni[i][j] + a.b_ih[512 + c] + r * (acc_nh[i][j] + a.b_hh[512 + c]));
            o[j] = (1.0f - z) * n + z * hp[j];
        }
        *reinterpret_cast<float4 *>(a.h_out + row * 256 + c0) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------- stage 2
// EdgeAttention (srnn_model.py:256-339) scores are q.k_i with q = W_t o_t + b_t and k_i = W_s o_i + b_s.  Since
//   q.(W_s o_i + b_s) = (W_s^T q).o_i + q.b_s   and the softmax over i ignores the per-env constant q.b_s,
// the N*H x 64 key projection is replaced by ONE per-env projection qt = (W_s^T W_t) o_t + W_s^T b_t (256 values)
// and the scores are taken against the edge outputs o_i that the weighted sum reads anyway.
__global__ void fold_attention_kernel(const float *__restrict__ wt, const float *__restrict__ bt, const float *__restrict__ wsp,
                                      float *__restrict__ wc, float *__restrict__ bc)
{
    const int a = blockIdx.x, b = threadIdx.x;       // wc[a][b] = sum_d wsp[d][a] * wt[d][b]; grid 256 x 256 threads
    double acc = 0.0;
    for (int d = 0; d < 64; ++d) acc += (double)wsp[d * 256 + a] * (double)wt[d * 256 + b];
    wc[a * 256 + b] = (float)acc;
    if (b == 0) {
        double s = 0.0;
        for (int d = 0; d < 64; ++d) s += (double)wsp[d * 256 + a] * (double)bt[d];
        bc[a] = (float)s;
    }
}

// one warp per env: ONE pass over the env's H edge outputs (each row is read from HBM once): score against qt,
// online softmax (running max / sum, rescaled accumulator), weighted sum; writes cat = [o_t | c]  (N x 512)
__global__ void __launch_bounds__(128) attention_kernel(const float *__restrict__ h_edge, const float *__restrict__ qt,
                                                        float *__restrict__ cat, int N, int H)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * 4 + warp;
    pdl_wait();                       // qt / h_edge come from the two kernels before this one
    pdl_launch_dependents();          // the node / heads kernel may set itself up beside this (bandwidth-bound) kernel
    if (e >= N) return;
    const float4 qa = *reinterpret_cast<const float4 *>(qt + (size_t)e * 256 + lane * 8);
    const float4 qb = *reinterpret_cast<const float4 *>(qt + (size_t)e * 256 + lane * 8 + 4);
    const float *ot = h_edge + (size_t)e * (H + 1) * 256;
    const float temp = (float)H / 8.0f;                       // num_edges / sqrt(attention_size)
    float run_max = -INFINITY, run_sum = 0.0f;
    float c[8] = {};
    float4 n0 = *reinterpret_cast<const float4 *>(ot + 256 + lane * 8), n1 = *reinterpret_cast<const float4 *>(ot + 256 + lane * 8 + 4);
    for (int i = 0; i < H; ++i) {
        const float4 v0 = n0, v1 = n1;
        if (i + 1 < H) {                                      // software prefetch of the next row
            const float *os = ot + (size_t)(2 + i) * 256 + lane * 8;
            n0 = *reinterpret_cast<const float4 *>(os); n1 = *reinterpret_cast<const float4 *>(os + 4);
        }
        float s = qa.x * v0.x + qa.y * v0.y + qa.z * v0.z + qa.w * v0.w + qb.x * v1.x + qb.y * v1.y + qb.z * v1.z + qb.w * v1.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s *= temp;
        const float new_max = fmaxf(run_max, s);
        const float scale = expf(run_max - new_max);          // 0 on the first row (run_max = -inf)
        const float p = expf(s - new_max);
        run_sum = run_sum * scale + p;
        c[0] = fmaf(p, v0.x, c[0] * scale); c[1] = fmaf(p, v0.y, c[1] * scale); c[2] = fmaf(p, v0.z, c[2] * scale); c[3] = fmaf(p, v0.w, c[3] * scale);
        c[4] = fmaf(p, v1.x, c[4] * scale); c[5] = fmaf(p, v1.y, c[5] * scale); c[6] = fmaf(p, v1.z, c[6] * scale); c[7] = fmaf(p, v1.w, c[7] * scale);
        run_max = new_max;
    }
    const float inv = 1.0f / run_sum;
    float *dst = cat + (size_t)e * 512;
    *reinterpret_cast<float4 *>(dst + lane * 8) = *reinterpret_cast<const float4 *>(ot + lane * 8);
    *reinterpret_cast<float4 *>(dst + lane * 8 + 4) = *reinterpret_cast<const float4 *>(ot + lane * 8 + 4);
    *reinterpret_cast<float4 *>(dst + 256 + lane * 8) = make_float4(c[0] * inv, c[1] * inv, c[2] * inv, c[3] * inv);
    *reinterpret_cast<float4 *>(dst + 256 + lane * 8 + 4) = make_float4(c[4] * inv, c[5] * inv, c[6] * inv, c[7] * inv);
}

// critic_linear (256 -> 1) and dist.fc_mean (256 -> 2): one warp per env, three dot products
__global__ void __launch_bounds__(128) heads_kernel(const float *__restrict__ c2, const float *__restrict__ feat,
                                                    const float *__restrict__ wv, const float *__restrict__ bv,
                                                    const float *__restrict__ wm, const float *__restrict__ bm,
                                                    float *__restrict__ value, float *__restrict__ mean, int N)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * 4 + warp;
    if (e >= N) return;
    float v = 0.f, m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int k = half * 128 + lane * 4;
        const float4 c = *reinterpret_cast<const float4 *>(c2 + (size_t)e * 256 + k);
        const float4 f = *reinterpret_cast<const float4 *>(feat + (size_t)e * 256 + k);
        const float4 a = *reinterpret_cast<const float4 *>(wv + k);
        const float4 b0 = *reinterpret_cast<const float4 *>(wm + k), b1 = *reinterpret_cast<const float4 *>(wm + 256 + k);
        v += c.x * a.x + c.y * a.y + c.z * a.z + c.w * a.w;
        m0 += f.x * b0.x + f.y * b0.y + f.z * b0.z + f.w * b0.w;
        m1 += f.x * b1.x + f.y * b1.y + f.z * b1.z + f.w * b1.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        m0 += __shfl_xor_sync(0xffffffffu, m0, o);
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
    }
    if (lane == 0) { value[e] = v + bv[0]; mean[2 * (size_t)e] = m0 + bm[0]; mean[2 * (size_t)e + 1] = m1 + bm[1]; }
}

// ---------------------------------------------------------------------------------------------- stage 3 helpers
// x[:, 0:64] = ReLU(W_ne (W_r robot_node + b_r) + b_ne)     (robot_linear 7->3, encoder_linear 3->64)
__global__ void node_encode_kernel(const float *__restrict__ robot_node, const float *__restrict__ rw, const float *__restrict__ rb,
                                   const float *__restrict__ ew, const float *__restrict__ eb, float *__restrict__ x, int N)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * 64) return;
    const int e = (int)(idx >> 6), k = (int)(idx & 63);
    const float *rn = robot_node + (size_t)e * 7;
    float r3[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 7; ++i) s = fmaf(rw[j * 7 + i], rn[i], s);
        r3[j] = s + rb[j];
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) s = fmaf(ew[k * 3 + j], r3[j], s);
    x[(size_t)e * 128 + k] = fmaxf(s + eb[k], 0.0f);
}

// node GRU gates: gi, gh [N,384] (biases already added), h_prev [N,128] masked -> h_out [N,128]
__global__ void node_gru_gate_kernel(const float *__restrict__ gi, const float *__restrict__ gh, const float *__restrict__ h_prev,
                                     const float *__restrict__ masks, float *__restrict__ h_out, int N)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * 128) return;
    const int e = (int)(idx >> 7), c = (int)(idx & 127);
    const float *i3 = gi + (size_t)e * 384, *h3 = gh + (size_t)e * 384;
    const float r = sigmoidf_(i3[c] + h3[c]);
    const float z = sigmoidf_(i3[128 + c] + h3[128 + c]);
    const float n = tanhf(i3[256 + c] + r * h3[256 + c]);
    h_out[idx] = (1.0f - z) * n + z * (h_prev[idx] * masks[e]);
}

// ---------------------------------------------------------------------------------------------- host side
struct Workspace {
    float *qt, *cat, *x, *gi, *gh, *y, *ac1, *a2, *c2;
};

static size_t carve_ws(Workspace *w, void *base, int N, int H)
{
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](float **p, size_t count) {
        if (w) *p = (float *)(b ? b + off : nullptr);
        off = (off + count * sizeof(float) + 255) & ~(size_t)255;
    };
    Workspace tmp;
    Workspace *q = w ? w : &tmp;
    (void)H;
    take(&q->qt, (size_t)N * 256);
    take(&q->cat, (size_t)N * 512);
    take(&q->x, (size_t)N * 128);
    take(&q->gi, (size_t)N * 384);
    take(&q->gh, (size_t)N * 384);
    take(&q->y, (size_t)N * 256);
    take(&q->ac1, (size_t)N * 512);     // [tanh(actor.0 y) | tanh(critic.0 y)]
    take(&q->a2, (size_t)N * 256);
    take(&q->c2, (size_t)N * 256);
    return off;
}

size_t dsrnn_workspace_bytes(int n_envs, int human_num) { return carve_ws(nullptr, nullptr, n_envs, human_num); }

static void destroy_tc_linears(CnDsrnn *m)
{
    TcLinear *all[] = {&m->att_qt, &m->emb, &m->gi, &m->gh, &m->out, &m->ac0, &m->actor2, &m->critic2};
    for (TcLinear *L : all) tc_linear_destroy(L);
    if (m->att_wc) cudaFree(m->att_wc);
    if (m->att_bc) cudaFree(m->att_bc);
    m->att_wc = m->att_bc = nullptr;
}

// `create`: allocate the packed images and pack; otherwise re-pack the current weights into the existing images
static const char *pack_tc_linears(CnDsrnn *m, cudaStream_t s, bool create)
{
    const CnDsrnnWeights &w = m->w;
    const char *msg;
#define CN_TCL(L, W0, B0, N0, W1, B1, N1, K, NT) \
    if ((msg = create ? tc_linear_create(&m->L, W0, B0, N0, W1, B1, N1, K, NT, s) : tc_linear_repack(&m->L, W0, B0, N0, W1, B1, N1, s))) return msg
    if (create && (cudaMalloc(&m->att_wc, 256 * 256 * sizeof(float)) != cudaSuccess || cudaMalloc(&m->att_bc, 256 * sizeof(float)) != cudaSuccess))
        return "cudaMalloc of the folded attention projection failed";
    fold_attention_kernel<<<256, 256, 0, s>>>(w.att_t_w, w.att_t_b, w.att_s_w, m->att_wc, m->att_bc);
    CN_TCL(att_qt, m->att_wc, m->att_bc, 256, nullptr, nullptr, 0, 256, 256);
    CN_TCL(emb, w.n_att_w, w.n_att_b, 64, nullptr, nullptr, 0, 512, 64);
    CN_TCL(gi, w.n_w_ih, w.n_b_ih, 384, nullptr, nullptr, 0, 128, 192);
    CN_TCL(gh, w.n_w_hh, w.n_b_hh, 384, nullptr, nullptr, 0, 128, 192);
    CN_TCL(out, w.n_out_w, w.n_out_b, 256, nullptr, nullptr, 0, 128, 256);
    CN_TCL(ac0, w.actor0_w, w.actor0_b, 256, w.critic0_w, w.critic0_b, 256, 256, 256);
    CN_TCL(actor2, w.actor2_w, w.actor2_b, 256, nullptr, nullptr, 0, 256, 256);
    CN_TCL(critic2, w.critic2_w, w.critic2_b, 256, nullptr, nullptr, 0, 256, 256);
#undef CN_TCL
    return cudaGetLastError() == cudaSuccess ? nullptr : "packing the linear layers failed";
}

const char *dsrnn_create(const CnDsrnnWeights *w, int device, cudaStream_t stream, CnDsrnn **out)
{
    CnDsrnn *m = new (std::nothrow) CnDsrnn();
    if (!m) return "out of host memory";
    m->w = *w;
    m->device = device;
    m->last_launches = 0;
    m->tc_state = nullptr;
    m->timing = false;
    cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device);
#ifdef CN_DEBUG_SWITCHES
    const char *env = getenv("CN_NODE_UNFUSED");      // development builds only: one launch per node / head layer
    m->unfused_node = env && env[0] == '1';
#endif
    const char *msg = dsrnn_tc_create(w, stream, &m->tc_state);
    if (!msg) msg = dsrnn_node_tc_create(w, stream, &m->node_state);
    if (!msg) msg = pack_tc_linears(m, stream, true);
    if (msg) { dsrnn_destroy(m); return msg; }
    *out = m;
    return nullptr;
}

void dsrnn_destroy(CnDsrnn *m)
{
    if (m->tc_state) dsrnn_tc_destroy(m->tc_state);
    if (m->node_state) dsrnn_node_tc_destroy(m->node_state);
    destroy_tc_linears(m);
    dsrnn_time_ms(m, nullptr);
    for (auto &p : m->pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    delete m;
}

// New weight VALUES (after an optimiser step or load_state_dict), possibly at new addresses: every packed image is
// re-written in place by its pack kernel on `stream`.  Nothing is freed or allocated and nothing synchronises, so the
// images keep their addresses -- CUDA graphs that captured a forward (rollout.GraphedRollout) stay valid and simply see
// the new weights -- and a failure leaves the previous state usable.
const char *dsrnn_update_weights(CnDsrnn *m, const CnDsrnnWeights *w, cudaStream_t stream)
{
    if (!m->tc_state || !m->node_state) return "dsrnn_update_weights: model was not fully created";
    m->w = *w;
    const char *msg = dsrnn_tc_repack(m->tc_state, w, stream);
    if (!msg) msg = dsrnn_node_tc_repack(m->node_state, w, stream);
    return msg ? msg : pack_tc_linears(m, stream, false);
}

int dsrnn_last_launches(const CnDsrnn *m) { return m->last_launches; }
int dsrnn_device(const CnDsrnn *m) { return m->device; }
const char *dsrnn_tc_edge_sequence_step(void *state, int n_envs, int H, const CnEdgeSeqStep *io, cudaStream_t stream);
const char *dsrnn_edge_sequence_step(CnDsrnn *m, int n_envs, int human_num, const CnEdgeSeqStep *io, cudaStream_t stream)
{
    return dsrnn_tc_edge_sequence_step(m->tc_state, n_envs, human_num, io, stream);
}
void dsrnn_set_refill_env(CnDsrnn *m, CnEnv *env) { m->refill_env = env; }
void dsrnn_set_edge_event(CnDsrnn *m, void *event) { m->edge_done = (cudaEvent_t)event; }
void dsrnn_set_edge_image(CnDsrnn *m, void *in_hi, void *in_lo, void *out_hi, void *out_lo)
{
    m->edge_img[0] = in_hi; m->edge_img[1] = in_lo; m->edge_img[2] = out_hi; m->edge_img[3] = out_lo;
}

// one linear layer, on CUDA cores (fp32) or tensor cores (bf16x3 / bf16)
struct LinearRun {
    CnDsrnn *m; int precision; cudaStream_t s; int *launches; const char *err = nullptr;
    void operator()(const TcLinear &tcl, const float *W, const float *b, TcLinearCall c, int N, int K)
    {
        if (err) return;
        if (precision == CN_PREC_FP32) {
            LinArgs a;
            a.X = c.X; a.ldx = c.ldx; a.rows_per_env = c.rows_per_env; a.env_stride_rows = c.env_stride_rows; a.first_row = c.first_row;
            a.rowscale = c.rowscale; a.W = W; a.b = b; a.M = c.M; a.N = N; a.K = K; a.Y = c.Y; a.ldy = c.ldy; a.ycol0 = c.ycol0; a.act = c.act;
            launch_linear(a, s, launches);
        } else {
            c.three_pass = precision == CN_PREC_BF16X3 ? 1 : 0;
            c.fp16 = precision == CN_PREC_FP16 ? 1 : 0;
            err = tc_linear_run(&tcl, c, m->num_sms, s);
            ++*launches;
        }
    }
};

static TcLinearCall call(const float *X, int ldx, int M, float *Y, int ldy, int act, int ycol0 = 0)
{
    TcLinearCall c;
    c.X = X; c.ldx = ldx; c.M = M; c.Y = Y; c.ldy = ldy; c.act = act; c.ycol0 = ycol0;
    return c;
}

const char *dsrnn_forward(CnDsrnn *m, int N, int H, const CnDsrnnIO *io, int precision, void *workspace, cudaStream_t s)
{
    const CnDsrnnWeights &w = m->w;
    Workspace ws;
    carve_ws(&ws, workspace, N, H);
    int launches = 0;

    // ---- stage 1: edge GRUs -> io->h_edge_out
    if (m->timing) {
        std::pair<cudaEvent_t, cudaEvent_t> p;
        if (!m->pool.empty()) { p = m->pool.back(); m->pool.pop_back(); }
        else { cudaEventCreate(&p.first); cudaEventCreate(&p.second); }
        cudaEventRecord(p.first, s);
        m->pending.push_back(p);
    }
    if (precision == CN_PREC_FP32) {
        EdgeArgs t;
        t.x_in = io->temporal_edges; t.h_in = io->h_edge_in; t.masks = io->masks;
        t.enc_w = w.t_enc_w; t.enc_b = w.t_enc_b; t.w_ih = w.t_w_ih; t.w_hh = w.t_w_hh; t.b_ih = w.t_b_ih; t.b_hh = w.t_b_hh;
        t.h_out = io->h_edge_out; t.M = N; t.rpe = 1; t.stride = H + 1; t.first = 0;
        edge_gru_simt_kernel<<<dim3((t.M + LBM - 1) / LBM, 256 / LBN), 256, 0, s>>>(t);
        EdgeArgs sp = t;
        sp.x_in = io->spatial_edges;
        sp.enc_w = w.s_enc_w; sp.enc_b = w.s_enc_b; sp.w_ih = w.s_w_ih; sp.w_hh = w.s_w_hh; sp.b_ih = w.s_b_ih; sp.b_hh = w.s_b_hh;
        sp.M = N * H; sp.rpe = H; sp.first = 1;
        edge_gru_simt_kernel<<<dim3((sp.M + LBM - 1) / LBM, 256 / LBN), 256, 0, s>>>(sp);
        launches += 2;
    } else {
        const char *msg = dsrnn_tc_edge_forward(m->tc_state, &w, N, H, io, precision, s, &launches, m->edge_img);
        if (msg) return msg;
    }
    if (m->timing) cudaEventRecord(m->pending.back().second, s);
    // the machine-filling part of the forward ends here; what follows (projection, attention, node / heads) leaves most SMs
    // idle, so a caller that pipelines half batches (rollout.PipelinedRollout) starts the other half's crowd step on this event
    if (m->edge_done && cudaEventRecord(m->edge_done, s) != cudaSuccess) return "cudaEventRecord (edge event) failed";

    LinearRun run{m, precision, s, &launches};

    // ---- stage 2: attention projections, softmax, weighted sum
    {
        TcLinearCall q = call(io->h_edge_out, 256, N, ws.qt, 256, ACT_NONE);
        q.rows_per_env = 1; q.env_stride_rows = H + 1; q.first_row = 0;
        run(m->att_qt, m->att_wc, m->att_bc, q, 256, 256);
        // the attention kernel is bandwidth-bound with a small footprint: the one place in this forward where the env's
        // spare-episode refill (crowd_reset.cu) can share the SMs instead of waiting for them or making others wait
        if (m->refill_env) {
            if (cn_env_refill(m->refill_env, s) != CN_OK) return "cn_env_refill (refill hook) failed";
            ++launches;
        }
        if (cn_launch(attention_kernel, dim3((N + 3) / 4), dim3(128), 0, s, CN_PDL_ATTENTION, (const float *)io->h_edge_out, (const float *)ws.qt, ws.cat, N, H) != cudaSuccess)
            return "attention_kernel launch failed";
        ++launches;
    }

    // ---- stages 3 + 4 on the tensor cores: one kernel for the node RNN and the heads
    if (precision != CN_PREC_FP32 && !m->unfused_node) {
        const char *msg = dsrnn_node_tc_forward(m->node_state, N, io, ws.cat, io->actor_features, precision, s, &launches);
        if (msg) return msg;
        if (run.err) return run.err;
        m->last_launches = launches;
        const cudaError_t err = cudaGetLastError();
        return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
    }

    // ---- stage 3: node RNN
    {
        node_encode_kernel<<<(unsigned)(((size_t)N * 64 + 255) / 256), 256, 0, s>>>(io->robot_node, w.robot_w, w.robot_b,
                                                                                  w.n_enc_w, w.n_enc_b, ws.x, N);
        ++launches;
        run(m->emb, w.n_att_w, w.n_att_b, call(ws.cat, 512, N, ws.x, 128, ACT_RELU, 64), 64, 512);
        run(m->gi, w.n_w_ih, w.n_b_ih, call(ws.x, 128, N, ws.gi, 384, ACT_NONE), 384, 128);
        TcLinearCall gh = call(io->h_node_in, 128, N, ws.gh, 384, ACT_NONE);
        gh.rowscale = io->masks;
        run(m->gh, w.n_w_hh, w.n_b_hh, gh, 384, 128);
        node_gru_gate_kernel<<<(unsigned)(((size_t)N * 128 + 255) / 256), 256, 0, s>>>(ws.gi, ws.gh, io->h_node_in, io->masks,
                                                                                    io->h_node_out, N);
        ++launches;
        run(m->out, w.n_out_w, w.n_out_b, call(io->h_node_out, 128, N, ws.y, 256, ACT_NONE), 256, 128);
    }

    // ---- stage 4: heads
    {
        float *feat = io->actor_features ? io->actor_features : ws.a2;
        if (precision == CN_PREC_FP32) {
            run(m->ac0, w.actor0_w, w.actor0_b, call(ws.y, 256, N, ws.ac1, 512, ACT_TANH, 0), 256, 256);
            run(m->ac0, w.critic0_w, w.critic0_b, call(ws.y, 256, N, ws.ac1, 512, ACT_TANH, 256), 256, 256);
        } else {
            run(m->ac0, nullptr, nullptr, call(ws.y, 256, N, ws.ac1, 512, ACT_TANH, 0), 512, 256);   // actor.0 and critic.0 in one GEMM
        }
        run(m->actor2, w.actor2_w, w.actor2_b, call(ws.ac1, 512, N, feat, 256, ACT_TANH), 256, 256);
        run(m->critic2, w.critic2_w, w.critic2_b, call(ws.ac1 + 256, 512, N, ws.c2, 256, ACT_TANH), 256, 256);
        if (run.err) return run.err;
        heads_kernel<<<(N + 3) / 4, 128, 0, s>>>(ws.c2, feat, w.critic_lin_w, w.critic_lin_b, w.mean_w, w.mean_b, io->value, io->action_mean, N);
        ++launches;
    }
    if (run.err) return run.err;
    m->last_launches = launches;
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}
