// dsrnn_tc_linear.cu -- tall-skinny linear layers of the DS-RNN forward on tcgen05 (sm_100a):
//     Y[m, ycol0 + n] = act( sum_k X[row(m), k] * W[n, k] + b[n] ),   M up to N_envs*H rows, K in {128,256,512}, N <= 512
// used for the attention projections (srnn_model.py:256-339), the node RNN (:149-173) and the actor/critic
// heads (:378-395,487-495) when the precision is bf16x3 / bf16.
//
// One persistent CTA per SM walks (128-row tile, n_tile column block) work items.  Both operands stream through
// 2-slot shared-memory rings, 64 k-columns at a time:
//   warps 0-7  read the fp32 activations (row gather + optional per-env mask), split them into bf16 hi/lo and
//              write the 128B-swizzled K-major A image of the k-block; after the last k-block they are the
//              epilogue (tcgen05.ld -> bias -> ReLU/tanh -> fp32 store);
//   warp 8     cp.async.bulk's the pre-swizzled weight image of the k-block (hi | lo, n_tile rows);
//   warp 9     owns TMEM and issues tcgen05.mma (M=128, N=n_tile, K=16), 3 passes for bf16x3.
#include <new>
#include "dsrnn.cuh"
#include "dsrnn_tc_linear.cuh"
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kRows = 128;
constexpr int kThreads = 320;              // 8 staging/epilogue warps + weight producer + MMA issuer
constexpr int kASlotBytes = 2 * kRows * 128;          // hi | lo images of one 128 x 64 k-block
constexpr int kMaxNTile = 256;
constexpr int kBSlotBytes = 2 * kMaxNTile * 128;      // hi | lo images of n_tile x 64 weights
constexpr int kOffA = 0;
constexpr int kOffB = kOffA + 2 * kASlotBytes;        //  65536
constexpr int kOffBias = kOffB + 2 * kBSlotBytes;     // 196608
constexpr int kOffBar = kOffBias + kMaxNTile * 4;     // 197632
constexpr int kSmemBytes = kOffBar + 128;

struct LinTcArgs {
    const float *X; int ldx;
    int rows_per_env, env_stride_rows, first_row;
    const float *rowscale;
    const __nv_bfloat16 *wimg;     // [n_blocks][k_blocks] x (hi | lo) images of n_tile x 64
    const float *bias;             // [n_blocks * n_tile] (zero padded) or NULL
    int M, N, K, n_tile, n_blocks, m_tiles;
    float *Y; int ldy, ycol0;
    int act, three_pass, fp16;
};

__global__ void __launch_bounds__(kThreads, 1) linear_tc_kernel(const __grid_constant__ LinTcArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    float *s_bias = reinterpret_cast<float *>(smem + kOffBias);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kOffBar + 96);
    // barrier map: 0,1 a_full | 2,3 a_empty | 4,5 b_full | 6,7 b_empty | 8 tmem_full | 9 tmem_empty
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = a.K >> 6;
    const int items = a.m_tiles * a.n_blocks;
    const uint32_t b_bytes = (uint32_t)(a.three_pass ? 2 : 1) * a.n_tile * 128;
    const size_t img_kb_bytes = (size_t)3 * a.n_tile * 128;      // per (n block, k block): bf16 hi | bf16 lo | fp16 images

    if (threadIdx.x == 0) {
        mbar_init(bar(0), 256); mbar_init(bar(1), 256);
        mbar_init(bar(2), 1); mbar_init(bar(3), 1);
        mbar_init(bar(4), 1); mbar_init(bar(5), 1);
        mbar_init(bar(6), 1); mbar_init(bar(7), 1);
        mbar_init(bar(8), 1); mbar_init(bar(9), 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();                       // everything above touched only this CTA's shared memory / TMEM
    pdl_launch_dependents();

    if (warp < 8) {
        // =============================================================== A staging + epilogue
        const int tid = (warp & 3) * 32 + lane;      // row of the tile (TMEM lane) owned in the epilogue
        const int chalf = warp >> 2;                 // warps 0-3 take the even 32-column blocks, warps 4-7 the odd ones
        const int sub = lane >> 4, c4 = (lane & 15) * 4;     // two rows per warp instruction, 16 lanes x float4 each
        uint32_t it = 0, item_iter = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_iter) {
            const int mt = item % a.m_tiles, nb = item / a.m_tiles;
            const int row0 = mt * kRows;
            // row pointers / masks of the 8 rows this lane stages are fixed for the item
            const float *rowp[8];
            float rsc[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int m = row0 + warp * 16 + b * 2 + sub;
                rowp[b] = nullptr; rsc[b] = 1.f;
                if (m < a.M) {
                    const int env = m / a.rows_per_env;
                    rowp[b] = a.X + ((size_t)env * a.env_stride_rows + a.first_row + (m - env * a.rows_per_env)) * a.ldx + c4;
                    if (a.rowscale) rsc[b] = a.rowscale[env];
                }
            }
            float4 cur[8], nxt[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) cur[b] = rowp[b] ? *reinterpret_cast<const float4 *>(rowp[b]) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                if (kb + 1 < kblocks) {      // the next k-block's loads are in flight while this one is converted
#pragma unroll
                    for (int b = 0; b < 8; ++b)
                        nxt[b] = rowp[b] ? *reinterpret_cast<const float4 *>(rowp[b] + (kb + 1) * 64) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const uint32_t slot = it & 1u;
                mbar_wait(bar(2 + slot), ((it >> 1) & 1u) ^ 1u);
                unsigned char *dst = smem + kOffA + slot * kASlotBytes;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int r = warp * 16 + b * 2 + sub;
                    const float sc = rsc[b];
                    uint2 hi, lo;
                    if (a.fp16) { hi.x = pack_half2(cur[b].x * sc, cur[b].y * sc); hi.y = pack_half2(cur[b].z * sc, cur[b].w * sc); lo.x = lo.y = 0u; }
                    else { split_bf16x2(cur[b].x * sc, cur[b].y * sc, hi.x, lo.x); split_bf16x2(cur[b].z * sc, cur[b].w * sc, hi.y, lo.y); }
                    const int off = sw128_offset(r, c4);
                    *reinterpret_cast<uint2 *>(dst + off) = hi;
                    if (a.three_pass) *reinterpret_cast<uint2 *>(dst + kRows * 128 + off) = lo;
                }
                fence_proxy_async();
                mbar_arrive(bar(0 + slot));
#pragma unroll
                for (int b = 0; b < 8; ++b) cur[b] = nxt[b];
            }
            // ---- epilogue.  tcgen05.ld gives one row per lane; the tile is transposed through shared memory (the A ring is
            //      idle once the accumulators are complete) so every global store instruction writes whole 128-byte lines.
            if (a.bias) for (int i = threadIdx.x; i < a.n_tile; i += 256) s_bias[i] = a.bias[nb * a.n_tile + i];
            asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 epilogue warps only
            mbar_wait(bar(8), item_iter & 1u);
            tc_fence_after();
            float *scratch = reinterpret_cast<float *>(smem + kOffA) + warp * (32 * 33);   // 32 rows x 32 cols, row stride 33
            const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            const int rq = lane >> 3, cq = (lane & 7) * 4;       // store mapping: 4 rows x 128 B per instruction
            for (int cb = chalf; cb * 32 < a.n_tile; cb += 2) {
                float acc[32];
                tmem_ld16(t0 + cb * 32, acc);
                tmem_ld16(t0 + cb * 32 + 16, acc + 16);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = acc[j] + (a.bias ? s_bias[cb * 32 + j] : 0.f);
                    if (a.act == 1) v = fmaxf(v, 0.f);
                    else if (a.act == 2) v = fast_tanh(v);
                    scratch[lane * 33 + j] = v;
                }
                __syncwarp();
                const int ncol = nb * a.n_tile + cb * 32 + cq;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rl = rq + 4 * j;                   // row inside the warp's 32-row slice
                    const int m = row0 + (warp & 3) * 32 + rl;
                    const float *sp = scratch + rl * 33 + cq;
                    const float4 v = make_float4(sp[0], sp[1], sp[2], sp[3]);
                    if (m < a.M) {
                        float *y = a.Y + (size_t)m * a.ldy + a.ycol0 + ncol;
                        if (ncol + 4 <= a.N && ((a.ldy | a.ycol0) & 3) == 0) *reinterpret_cast<float4 *>(y) = v;
                        else {
                            if (ncol + 0 < a.N) y[0] = v.x;
                            if (ncol + 1 < a.N) y[1] = v.y;
                            if (ncol + 2 < a.N) y[2] = v.z;
                            if (ncol + 3 < a.N) y[3] = v.w;
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(bar(9));
            asm volatile("bar.sync 1, 256;" ::: "memory");        // scratch (= A ring) and s_bias are rewritten by the next item
        }
    } else if (warp == 8) {
        // =============================================================== weight producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int nb = item / a.m_tiles;
                const char *img = reinterpret_cast<const char *>(a.wimg) + (size_t)nb * kblocks * img_kb_bytes + (a.fp16 ? (size_t)2 * a.n_tile * 128 : 0);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t slot = it & 1u;
                    mbar_wait(bar(6 + slot), ((it >> 1) & 1u) ^ 1u);
                    mbar_expect_tx(bar(4 + slot), b_bytes);
                    bulk_g2s(s_base + kOffB + slot * kBSlotBytes, img + (size_t)kb * img_kb_bytes, b_bytes, bar(4 + slot));
                }
            }
        }
    } else {
        // =============================================================== MMA issuer: warp-uniform control flow, the elected lane issues
        {
            const uint32_t lead = elect_one_sync();
            uint32_t it = 0, item_iter = 0;
            const uint32_t idesc = a.fp16 ? idesc_f16(a.n_tile) : idesc_bf16(a.n_tile);
            // descriptor high word is constant; K steps / slots / hi-lo images only add to the low word (address >> 4 | LBO)
            constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            auto make_desc = [](uint32_t lo) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi)); return d; };
            const uint32_t desc_lo0 = (s_base >> 4) | (1u << 16);
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_iter) {
                mbar_wait(bar(9), (item_iter & 1u) ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t slot = it & 1u;
                    mbar_wait(bar(0 + slot), (it >> 1) & 1u);
                    mbar_wait(bar(4 + slot), (it >> 1) & 1u);
                    tc_fence_after();
                    const uint32_t a_hi = desc_lo0 + ((kOffA + slot * kASlotBytes) >> 4), a_lo = a_hi + ((kRows * 128) >> 4);
                    const uint32_t b_hi = desc_lo0 + ((kOffB + slot * kBSlotBytes) >> 4), b_lo = b_hi + ((a.n_tile * 128) >> 4);
                    const int passes = a.three_pass ? 3 : 1;
                    for (int ps = 0; ps < passes; ++ps) {
                        const uint32_t aa = ps == 1 ? a_lo : a_hi, bb = ps == 2 ? b_lo : b_hi;
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16)
                            umma_bf16_if(lead, tmem_base, make_desc(aa + 2 * k16), make_desc(bb + 2 * k16), idesc, (kb | ps | k16) != 0 ? 1u : 0u);
                    }
                    umma_commit_if(lead, bar(2 + slot));
                    umma_commit_if(lead, bar(6 + slot));
                }
                umma_commit_if(lead, bar(8));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// W = [W0 (n0 rows) ; W1 (n1 rows)] (row-major [*, K]) -> per (n block, k block): hi image | lo image, zero padded
__global__ void pack_linear_kernel(const float *W0, int n0, const float *W1, int n1, const float *b0, const float *b1,
                                   int K, int n_tile, int n_blocks, __nv_bfloat16 *wimg, float *bias)
{
    const int kblocks = K >> 6;
    const long total = (long)n_blocks * kblocks * n_tile * 64;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long t = idx;
        const int k = (int)(t & 63); t >>= 6;
        const int q = (int)(t % n_tile); t /= n_tile;
        const int kb = (int)(t % kblocks); t /= kblocks;
        const int nb = (int)t;
        const int n = nb * n_tile + q;
        float w = 0.f;
        if (n < n0) w = W0[(size_t)n * K + kb * 64 + k];
        else if (n < n0 + n1) w = W1[(size_t)(n - n0) * K + kb * 64 + k];
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        char *base = reinterpret_cast<char *>(wimg) + ((size_t)nb * kblocks + kb) * 3 * n_tile * 128;
        *reinterpret_cast<__nv_bfloat16 *>(base + sw128_offset(q, k)) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(base + (size_t)n_tile * 128 + sw128_offset(q, k)) = lo;
        *reinterpret_cast<__half *>(base + (size_t)2 * n_tile * 128 + sw128_offset(q, k)) = __float2half_rn(w);
    }
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_blocks * n_tile; n += gridDim.x * blockDim.x)
        bias[n] = n < n0 ? (b0 ? b0[n] : 0.f) : (n < n0 + n1 ? (b1 ? b1[n - n0] : 0.f) : 0.f);
}

}  // namespace

const char *tc_linear_create(TcLinear *L, const float *W0, const float *b0, int n0, const float *W1, const float *b1, int n1,
                             int K, int n_tile, cudaStream_t stream)
{
    if (K % 64 != 0 || n_tile % 16 != 0 || n_tile < 16 || n_tile > kMaxNTile) return "tc_linear_create: unsupported shape";
    L->N = n0 + n1; L->K = K; L->n_tile = n_tile;
    L->n_blocks = (L->N + n_tile - 1) / n_tile;
    const size_t bytes = (size_t)L->n_blocks * (K / 64) * 3 * n_tile * 128;
    if (cudaMalloc(&L->wimg, bytes) != cudaSuccess || cudaMalloc(&L->bias, (size_t)L->n_blocks * n_tile * sizeof(float)) != cudaSuccess)
        return "tc_linear_create: cudaMalloc failed";
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024) != cudaSuccess)
            return "tc_linear_create: cudaFuncSetAttribute failed";
        attr_set = true;
    }
    return tc_linear_repack(L, W0, b0, n0, W1, b1, n1, stream);
}

// packs new values of the same weights into the existing images (stable addresses, no allocation, no synchronisation)
const char *tc_linear_repack(TcLinear *L, const float *W0, const float *b0, int n0, const float *W1, const float *b1, int n1,
                             cudaStream_t stream)
{
    if (!L->wimg || n0 + n1 != L->N) return "tc_linear_repack: layer was not created with this shape";
    pack_linear_kernel<<<128, 256, 0, stream>>>(W0, n0, W1, n1, b0, b1, L->K, L->n_tile, L->n_blocks,
                                                reinterpret_cast<__nv_bfloat16 *>(L->wimg), L->bias);
    return cudaGetLastError() == cudaSuccess ? nullptr : "pack_linear_kernel launch failed";
}

void tc_linear_destroy(TcLinear *L)
{
    if (L->wimg) cudaFree(L->wimg);
    if (L->bias) cudaFree(L->bias);
    L->wimg = nullptr; L->bias = nullptr;
}

const char *tc_linear_run(const TcLinear *L, const TcLinearCall &c, int num_sms, cudaStream_t stream)
{
    LinTcArgs a;
    a.X = c.X; a.ldx = c.ldx; a.rows_per_env = c.rows_per_env; a.env_stride_rows = c.env_stride_rows; a.first_row = c.first_row;
    a.rowscale = c.rowscale;
    a.wimg = reinterpret_cast<const __nv_bfloat16 *>(L->wimg); a.bias = L->bias;
    a.M = c.M; a.N = L->N; a.K = L->K; a.n_tile = L->n_tile; a.n_blocks = L->n_blocks;
    a.m_tiles = (c.M + kRows - 1) / kRows;
    a.Y = c.Y; a.ldy = c.ldy; a.ycol0 = c.ycol0; a.act = c.act; a.three_pass = c.fp16 ? 0 : c.three_pass; a.fp16 = c.fp16;
    const int items = a.m_tiles * a.n_blocks;
    const cudaError_t lerr = cn_launch(linear_tc_kernel, dim3(items < num_sms ? items : num_sms), dim3(kThreads), kSmemBytes + 1024, stream, CN_PDL_LINEAR, a);
    if (lerr != cudaSuccess) return cudaGetErrorString(lerr);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}
