// gemm_bf16x3.cu -- the GEMM of the PPO update on the 5th-gen tensor cores (tcgen05 / TMEM / TMA),
// sm_100a only.  Replaces every cuBLAS product of the reference's update (ppo.py:36-118 -> evaluate_actions,
// srnn_model.py:53-104 and its autograd transpose): forward linears, the recurrent products of the masked GRU sequences
// and their backward, and the weight-gradient reductions over the T*R row dimension.
//
//     C[M,N] (fp32)   =  | +=  | atomically +=     (A_hi + A_lo)(B_hi + B_lo)   (+ bias[n], ReLU / tanh)
//
// Split-bf16 3-pass arithmetic (A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation in TMEM; operand error ~2^-16: the
// precision of the rollout kernels, CN_PREC_BF16X3).  Operands are bf16 (hi, lo) pairs in global memory, written by the
// kernels that produce them (cn_split_bf16, the gate kernels, the edge sequence kernel).  Each operand is either
//   K-major  : a row-major [MN, K] array (activations x weights^T: A = X [M,K], B = W [N,K]), or
//   MN-major : a row-major [K, MN] array -- what the backward products read without any transposition:
//              dX = dY W  (B = W [K=N_out, N=K_in] row-major), dW = dY^T X (A = dY [K=rows, M], B = X [K=rows, N]).
// Both forms are fetched by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) straight into the canonical UMMA shared-memory
// layouts; the MMA instruction descriptor carries the a_major / b_major bits.
//
// One persistent CTA per SM walks work items (problem, k-slice, 128-row tile, n-tile).  192 threads:
//   warp 0      TMA producer: one elected lane streams 64-deep k-blocks of A (hi, lo) and B (hi, lo) through a ring of stages;
//   warp 1      owns TMEM (512 columns = two accumulator buffers of up to 256 columns) and issues tcgen05.mma
//               (M = 128, N = n-tile, K = 16; 12 per k-block for the three passes);
//   warps 2-5   epilogue: tcgen05.ld -> (bias, activation) -> transposed through shared memory -> coalesced float4 stores,
//               read-modify-write (accumulate) or red.global.add (split-K partial sums), overlapped with the MMAs of the
//               next item through the second TMEM buffer.
// Up to kMaxProblems independent problems share one launch (grouped GEMM: the spatial / temporal edge GRUs have different
// weights; the weight gradients of both are one launch).
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include "tc_common.cuh"
#include "../../include/crowdnav_b200.h"

using namespace tc;

namespace {

constexpr int kMaxProblems = CN_GEMM_MAX_PROBLEMS;
constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;              // one 128 x 64 bf16 image (hi or lo) of A: 16 KB in either major
constexpr int kMaxStages = 6;
constexpr int kScratchBytes = 4 * 32 * 33 * 4;     // epilogue transposition: 4 warps x 32 rows x 33 floats
constexpr int kSmemLimit = 232448 - 1024;          // dynamic shared memory of one sm_100 CTA minus the alignment slack

struct alignas(64) GemmProblem {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    float *c;
    const float *bias;
    long long ldc;
    int m, n, k;
    int bn;                    // n-tile = MMA N (multiple of 16, <= 256)
    int m_tiles, n_tiles;
    int split_k, kb_per_split; // k-blocks (64 deep) per k-slice
    int a_mn, b_mn;            // operand majors (1 = MN-major)
    int a_lo_on, b_lo_on;      // passes A_lo*B_hi / A_hi*B_lo enabled
    int act, mode;             // mode 0 store, 1 read-modify-write add, 2 atomic add
    int item0, items;          // this problem's range of global work items
    unsigned stage_tx;         // bytes one stage of this problem brings in
};

struct GemmArgs {
    GemmProblem p[kMaxProblems];
    int n_problems, total_items;
    int stages, stage_bytes, b_off_lo;   // ring geometry: A_hi @0, A_lo @16K, B_hi @32K, B_lo @32K + b_off_lo
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t mbar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(mbar), "r"(c0), "r"(c1) : "memory");
}

struct Item { int p, ks, mt, nt; };

__device__ __forceinline__ Item decode_item(const GemmArgs &a, int item)
{
    Item it;
    int p = 0;
    while (p + 1 < a.n_problems && item >= a.p[p + 1].item0) ++p;
    const GemmProblem &g = a.p[p];
    int r = item - g.item0;
    it.p = p;
    it.nt = r % g.n_tiles; r /= g.n_tiles;      // n-tiles of the same rows run on neighbouring CTAs (A comes out of L2),
    it.mt = r % g.m_tiles; r /= g.m_tiles;      // then the m-tiles of one k-slice (split-K: B comes out of L2)
    it.ks = r;
    return it;
}

__global__ void __launch_bounds__(kThreads, 1) gemm_bf16x3_kernel(const __grid_constant__ GemmArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    const int ring_bytes = a.stages * a.stage_bytes;
    float *scratch_all = reinterpret_cast<float *>(smem + ring_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + ring_bytes + kScratchBytes);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 4);
    const uint32_t bar0 = smem_u32(bars);
    // barrier map: [0, S) full | [S, 2S) empty | 2S, 2S+1 tmem_full | 2S+2, 2S+3 tmem_empty      (S = kMaxStages)
    auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (uint32_t)(kMaxStages + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (uint32_t)(2 * kMaxStages + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (uint32_t)(2 * kMaxStages + 2 + b); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_tfull(0), 1); mbar_init(bar_tfull(1), 1);
        mbar_init(bar_tempty(0), 4); mbar_init(bar_tempty(1), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // =============================================================== TMA producer
        if (lane == 0) {
            uint32_t stage = 0, parity = 1;           // a fresh barrier passes a wait on parity 1
            for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
                const Item it = decode_item(a, item);
                const GemmProblem &g = a.p[it.p];
                const int kb0 = it.ks * g.kb_per_split;
                int kb1 = kb0 + g.kb_per_split;
                const int kb_total = (g.k + 63) >> 6;
                if (kb1 > kb_total) kb1 = kb_total;
                const int m0 = it.mt * kTileM, n0 = it.nt * g.bn;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_empty(stage), parity);
                    const uint32_t full = bar_full(stage);
                    mbar_expect_tx(full, g.stage_tx);
                    const uint32_t sa = s_base + stage * a.stage_bytes, sb = sa + 2 * kABytes;
                    const int k0 = kb << 6;
                    if (g.a_mn) {                      // [K, M] array: two 64-wide M atoms of 64 k-rows each
                        tma_load_2d(sa, &g.a_hi, m0, k0, full);
                        tma_load_2d(sa + 8192, &g.a_hi, m0 + 64, k0, full);
                        if (g.a_lo_on) { tma_load_2d(sa + kABytes, &g.a_lo, m0, k0, full); tma_load_2d(sa + kABytes + 8192, &g.a_lo, m0 + 64, k0, full); }
                    } else {                           // [M, K] array: one box of 128 rows x 64 k
                        tma_load_2d(sa, &g.a_hi, k0, m0, full);
                        if (g.a_lo_on) tma_load_2d(sa + kABytes, &g.a_lo, k0, m0, full);
                    }
                    if (g.b_mn) {
                        for (int j = 0; j * 64 < g.bn; ++j) {
                            tma_load_2d(sb + j * 8192, &g.b_hi, n0 + j * 64, k0, full);
                            if (g.b_lo_on) tma_load_2d(sb + a.b_off_lo + j * 8192, &g.b_lo, n0 + j * 64, k0, full);
                        }
                    } else {
                        tma_load_2d(sb, &g.b_hi, k0, n0, full);
                        if (g.b_lo_on) tma_load_2d(sb + a.b_off_lo, &g.b_lo, k0, n0, full);
                    }
                    if (++stage == (uint32_t)a.stages) { stage = 0; parity ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);     // SBO = 1024 B, version 1, SWIZZLE_128B
            auto make_desc = [](uint32_t lo) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi)); return d; };
            uint32_t stage = 0, parity = 0, iter = 0;
            for (int item = blockIdx.x; item < a.total_items; item += gridDim.x, ++iter) {
                const Item it = decode_item(a, item);
                const GemmProblem &g = a.p[it.p];
                const int kb0 = it.ks * g.kb_per_split;
                int kb1 = kb0 + g.kb_per_split;
                const int kb_total = (g.k + 63) >> 6;
                if (kb1 > kb_total) kb1 = kb_total;
                const uint32_t buf = iter & 1u;
                mbar_wait(bar_tempty(buf), ((iter >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * 256u;
                const uint32_t idesc = idesc_bf16(g.bn) | ((uint32_t)g.a_mn << 15) | ((uint32_t)g.b_mn << 16);
                // descriptor low words: (address >> 4) | LBO << 16.  K-major: LBO unused (1), a K = 16 step is 32 bytes along the
                // 128-byte row.  MN-major: LBO = 8192 B (the next 64-wide MN atom), a K = 16 step is 16 rows of 128 bytes.
                const uint32_t a_lbo = g.a_mn ? (8192u >> 4) << 16 : 1u << 16, b_lbo = g.b_mn ? (8192u >> 4) << 16 : 1u << 16;
                const uint32_t a_step = g.a_mn ? 2048u >> 4 : 32u >> 4, b_step = g.b_mn ? 2048u >> 4 : 32u >> 4;
                uint32_t first = 1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_full(stage), parity);
                    tc_fence_after();
                    const uint32_t sa = s_base + stage * a.stage_bytes, sb = sa + 2 * kABytes;
                    const uint32_t a_hi = (sa >> 4) | a_lbo, a_lo = ((sa + kABytes) >> 4) | a_lbo;
                    const uint32_t b_hi = (sb >> 4) | b_lbo, b_lo = ((sb + a.b_off_lo) >> 4) | b_lbo;
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16) {
                        umma_bf16(d, make_desc(a_hi + k16 * a_step), make_desc(b_hi + k16 * b_step), idesc, first ? 0u : 1u);
                        first = 0;
                    }
                    if (g.a_lo_on) {
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) umma_bf16(d, make_desc(a_lo + k16 * a_step), make_desc(b_hi + k16 * b_step), idesc, 1u);
                    }
                    if (g.b_lo_on) {
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) umma_bf16(d, make_desc(a_hi + k16 * a_step), make_desc(b_lo + k16 * b_step), idesc, 1u);
                    }
                    umma_commit(bar_empty(stage));
                    if (++stage == (uint32_t)a.stages) { stage = 0; parity ^= 1u; }
                }
                umma_commit(bar_tfull(buf));
            }
        }
    } else {
        // =============================================================== epilogue (warps 2-5; TMEM lane quarter = warp % 4)
        const int q = warp & 3;
        float *scratch = scratch_all + (warp - 2) * (32 * 33);
        const int rq = lane >> 3, cq = (lane & 7) * 4;       // store mapping: 4 rows x 128 B per instruction
        uint32_t iter = 0;
        for (int item = blockIdx.x; item < a.total_items; item += gridDim.x, ++iter) {
            const Item it = decode_item(a, item);
            const GemmProblem &g = a.p[it.p];
            const uint32_t buf = iter & 1u;
            const int kb0 = it.ks * g.kb_per_split;
            const bool empty_slice = kb0 >= ((g.k + 63) >> 6);        // never produced by the host's split; guards a zero-trip MMA loop
            mbar_wait(bar_tfull(buf), (iter >> 1) & 1u);
            tc_fence_after();
            const uint32_t t0 = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
            const int row0 = it.mt * kTileM + q * 32;
            const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.c) & 15) == 0);
            for (int cb = 0; cb * 32 < g.bn; ++cb) {
                float acc[32];
                tmem_ld16(t0 + cb * 32, acc);
                if (cb * 32 + 16 < g.bn) tmem_ld16(t0 + cb * 32 + 16, acc + 16);
                tmem_ld_wait();
                const int nbase = it.nt * g.bn + cb * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = empty_slice ? 0.f : acc[j];
                    if (g.mode != 2) {
                        if (g.bias && nbase + j < g.n) v += __ldg(g.bias + nbase + j);
                        if (g.act == 1) v = fmaxf(v, 0.f);
                        else if (g.act == 2) v = tanhf(v);
                    }
                    scratch[lane * 33 + j] = v;
                }
                __syncwarp();
                const int ncol = nbase + cq;
                const bool col_in_tile = cb * 32 + cq < g.bn;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rl = rq + 4 * j;
                    const int m = row0 + rl;
                    const float *sp = scratch + rl * 33 + cq;
                    float4 v = make_float4(sp[0], sp[1], sp[2], sp[3]);
                    if (m < g.m && col_in_tile && ncol < g.n) {
                        float *y = g.c + (size_t)m * g.ldc + ncol;
                        if (g.mode == 2) {
                            atomicAdd(y, v.x);
                            if (ncol + 1 < g.n) atomicAdd(y + 1, v.y);
                            if (ncol + 2 < g.n) atomicAdd(y + 2, v.z);
                            if (ncol + 3 < g.n) atomicAdd(y + 3, v.w);
                        } else if (vec_ok && ncol + 4 <= g.n) {
                            if (g.mode == 1) { const float4 o = *reinterpret_cast<const float4 *>(y); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                            *reinterpret_cast<float4 *>(y) = v;
                        } else {
                            const float vv[4] = {v.x, v.y, v.z, v.w};
                            for (int e = 0; e < 4; ++e)
                                if (ncol + e < g.n) y[e] = g.mode == 1 ? y[e] + vv[e] : vv[e];
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(buf));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encode_fn()
{
    static EncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess) fn = reinterpret_cast<EncodeTiled>(p);
    }
    return fn;
}

// tensor map of a row-major bf16 [rows, cols] array (ld elements between rows), box = box_rows x 64 columns, SWIZZLE_128B
bool make_map(CUtensorMap *map, const void *base, long long rows, long long cols, long long ld, int box_rows)
{
    EncodeTiled fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// returns NULL on success or a static error string (called from c_abi.cu)
const char *gemm_bf16x3_launch(const CnGemm *problems, int n_problems, cudaStream_t stream, int *items_out)
{
    static int num_sms = 0;
    static bool attr_set = false;
    if (n_problems < 1 || n_problems > kMaxProblems) return "n_problems out of range";
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit + 1024) != cudaSuccess)
            return "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
        attr_set = true;
    }
    static thread_local GemmArgs args;
    int max_b_bytes = 0;
    long long total_kb_items = 0;
    for (int i = 0; i < n_problems; ++i) {
        const CnGemm &q = problems[i];
        GemmProblem &g = args.p[i];
        if (q.m < 1 || q.n < 1 || q.k < 1) return "m, n, k must be positive";
        if (!q.a.hi || !q.b.hi || !q.c) return "operand hi pointers and c are required";
        if ((q.a.ld & 7) || (q.b.ld & 7)) return "operand leading dimensions must be multiples of 8 elements";
        if ((reinterpret_cast<uintptr_t>(q.a.hi) | reinterpret_cast<uintptr_t>(q.a.lo) | reinterpret_cast<uintptr_t>(q.b.hi) |
             reinterpret_cast<uintptr_t>(q.b.lo)) & 15) return "operand pointers must be 16-byte aligned";
        if (q.act < 0 || q.act > 2) return "unknown activation";
        int bn = q.n <= 256 ? ((q.n + 15) & ~15) : 256;
        if (q.n > 256) {       // even n-tiles (e.g. 320 -> 2 x 160) instead of a full tile and a sliver
            const int tiles = (q.n + 255) / 256;
            bn = (((q.n + tiles - 1) / tiles) + 15) & ~15;
        }
        g.bn = bn;
        g.m = q.m; g.n = q.n; g.k = q.k;
        g.m_tiles = (q.m + kTileM - 1) / kTileM;
        g.n_tiles = (q.n + bn - 1) / bn;
        g.a_mn = q.a.mn_major ? 1 : 0; g.b_mn = q.b.mn_major ? 1 : 0;
        g.a_lo_on = q.a.lo ? 1 : 0; g.b_lo_on = q.b.lo ? 1 : 0;
        g.c = q.c; g.ldc = q.ldc; g.bias = q.bias; g.act = q.act;
        const int kb_total = (q.k + 63) >> 6;
        int split = q.split_k;
        if (split < 0) return "split_k must be >= 0";
        if (split > kb_total) split = kb_total;
        g.split_k = split;                                   // 0 = decided below (auto), 1 = no split
        g.mode = split == 1 ? (q.accumulate ? 1 : 0) : 2;
        if (g.mode == 2 && (q.bias || q.act)) return "split-K products cannot carry a bias or an activation";
        const int b_tile = g.b_mn ? ((bn + 63) / 64) * 8192 : bn * 128;
        g.stage_tx = (unsigned)(kABytes * (1 + g.a_lo_on) + b_tile * (1 + g.b_lo_on));
        if (b_tile > max_b_bytes) max_b_bytes = b_tile;
        total_kb_items += (long long)kb_total * g.m_tiles * g.n_tiles;
        // A: K-major [m, k] box 128 rows; MN-major [k, m] box 64 k-rows.  B likewise with bn rows / 64 k-rows.
        const bool ok = (g.a_mn ? make_map(&g.a_hi, q.a.hi, q.k, q.m, q.a.ld, 64) : make_map(&g.a_hi, q.a.hi, q.m, q.k, q.a.ld, kTileM)) &&
                        (!q.a.lo || (g.a_mn ? make_map(&g.a_lo, q.a.lo, q.k, q.m, q.a.ld, 64) : make_map(&g.a_lo, q.a.lo, q.m, q.k, q.a.ld, kTileM))) &&
                        (g.b_mn ? make_map(&g.b_hi, q.b.hi, q.k, q.n, q.b.ld, 64) : make_map(&g.b_hi, q.b.hi, q.n, q.k, q.b.ld, bn)) &&
                        (!q.b.lo || (g.b_mn ? make_map(&g.b_lo, q.b.lo, q.k, q.n, q.b.ld, 64) : make_map(&g.b_lo, q.b.lo, q.n, q.k, q.b.ld, bn)));
        if (!ok) return "cuTensorMapEncodeTiled failed (driver entry point missing, or an operand it cannot describe)";
    }
    // split-K "auto" (0): cut the k range so that the launch has about two items per SM, shared out in proportion to the work
    int item0 = 0;
    for (int i = 0; i < n_problems; ++i) {
        GemmProblem &g = args.p[i];
        const int kb_total = (g.k + 63) >> 6;
        if (g.split_k == 0) {
            const long long tiles = (long long)g.m_tiles * g.n_tiles;
            const double share = (double)kb_total * tiles / (double)total_kb_items;
            int split = (int)(share * 2.0 * num_sms / (double)tiles + 0.5);
            if (split < 1) split = 1;
            if (split > kb_total) split = kb_total;
            g.split_k = split;
        }
        g.kb_per_split = (kb_total + g.split_k - 1) / g.split_k;
        g.split_k = (kb_total + g.kb_per_split - 1) / g.kb_per_split;     // no empty slices
        g.item0 = item0;
        g.items = g.split_k * g.m_tiles * g.n_tiles;
        item0 += g.items;
    }
    args.n_problems = n_problems;
    args.total_items = item0;
    args.b_off_lo = max_b_bytes;
    args.stage_bytes = 2 * kABytes + 2 * max_b_bytes;
    int stages = (kSmemLimit - kScratchBytes - 256) / args.stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return "tile does not fit two pipeline stages";
    args.stages = stages;
    const int smem = stages * args.stage_bytes + kScratchBytes + 256 + 1024;
    const int grid = args.total_items < num_sms ? args.total_items : num_sms;
    gemm_bf16x3_kernel<<<grid, kThreads, smem, stream>>>(args);
    if (items_out) *items_out = args.total_items;
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}
