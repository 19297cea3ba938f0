// dsrnn_edge_tc.cu -- K3 stage 1 on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
// The edge GRUs of the DS-RNN (HumanHumanEdgeRNN.forward, srnn_model.py:201-215; one GRU(64->256) step on
// N temporal + N*H spatial edge rows) are one tall GEMM per weight set:
//     [e | m*h]  (rows x 320)   x   [W_ih | W_hh]^T  (320 x 768)      e = ReLU(W_enc x + b)  (K = 2, computed in place)
// followed by the gate non-linearities.  Persistent CTA PAIRS (cluster of 2, tcgen05.mma.cta_group::2: M = 256 = 128 rows per
// CTA, the weight operand split between the two CTAs) walk 256-row tile pairs; 20 warps per CTA, setmaxnreg budgets:
//   warps 0-7   epilogue: tcgen05.ld the accumulators of a 64-hidden-unit column tile (n_i | r | z | n_h, 256 TMEM columns,
//               double-buffered), h_prev back out of the staged A image, gates with 5 MUFU per element, 256-bit stores of h',
//               tcgen05.st zeroes n_h for the next column tile.
//   warps 8-15  stage the tile's A operand straight from the fp32 hidden state in HBM (masking, fp32 -> split bf16 hi/lo,
//               128B-swizzled K-major smem image, 160 KB).  The loads of the next tile are issued before waiting for the MMAs
//               of the current one; the image is handed over in three pieces (k-block 0 | 1-2 | 3-4).
//   warp 16     streams this CTA's half (96 of 192 rows, 12 KB) of the pre-swizzled weight chunks with
//               cp.async.bulk.tensor .cta_group::2 into a 5-slot ring; both CTAs' copies complete on the leader's mbarrier.
//   warp 17     (leader CTA) issues every tcgen05.mma (N = 192, K = 16) and the multicast tcgen05.commit's; TMEM (512 columns)
//               is allocated for the pair by this warp of both CTAs.  The WHOLE warp runs the issue loop and the elect.sync lane
//               executes the instructions (warp-uniform control flow keeps descriptors in uniform registers: tc_common.cuh).
//   warp 18     resident-image instantiation only: issues the TMA loads of the A image (k-blocks 1-4) as the MMAs release it.
// Precision: CN_PREC_BF16X3 runs A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (fp32 accumulate in TMEM, ~2^-16 relative operand
// error); CN_PREC_BF16 / CN_PREC_FP16 run one pass.
// Development aids: -DEDGE_PROFILE adds per-role clock64 counters (tools/edge_profile.py) and CN_EDGE_DEBUG what-if switches.
#include <cuda.h>
#include <cstdlib>
#include <new>
#include "dsrnn.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;                 // rows per CTA; the CTA pair multiplies 256 rows per MMA (cta_group::2)
constexpr int kKBlocks = 5;                // 64-wide k-blocks: 1 encoded-input block + 4 hidden blocks
constexpr int kColTiles = 4;               // 64 hidden units per column tile
constexpr int kABlockBytes = kRows * 128;  // 128 rows x 64 bf16
constexpr int kBChunkRows = 192;                 // weight rows per (column tile, k-block, part): [n_i|r|z] or [r|z|n_h] x 64
constexpr int kBChunkBytes = kBChunkRows * 128;
constexpr int kBHalfBytes = kBChunkBytes / 2;    // each CTA of the pair holds 96 of the 192 weight rows (N is split)
constexpr int kBSlots = 5;
constexpr int kEpiWarps = 8, kStageWarps = 8, kWarpProducer = 16, kWarpMma = 17;
constexpr int kThreads = 20 * 32;            // 5 warpgroups: 2 epilogue, 2 staging, 1 (producer, MMA, 2 idle warps)
constexpr int kRegsEpi = 128, kRegsStage = 96, kRegsMisc = 32;   // setmaxnreg budgets (640 threads start at 96 registers each)
static_assert(256 * kRegsEpi + 256 * kRegsStage + 128 * kRegsMisc <= 640 * 96, "setmaxnreg only moves registers inside the CTA's launch allocation");
// mbarrier map
constexpr int kBarFull = 0, kBarEmpty = kBSlots, kBarAReady = 2 * kBSlots, kBarAFree = kBarAReady + 3,
              kBarTmemFull = kBarAFree + 3, kBarTmemEmpty = kBarTmemFull + 2, kBarHprev = kBarTmemEmpty + 2,
              kBarALoaded = kBarHprev + 1,      // resident-image path: TMA of k-blocks 1-2 / 3-4 of this CTA's rows has landed
              kNumBars = kBarALoaded + 2;
constexpr int kWarpAProducer = 18;             // resident-image path: issues the TMA loads of the A image

constexpr int kOffAHi = 0;
constexpr int kOffALo = kOffAHi + kKBlocks * kABlockBytes;        //  81920
constexpr int kOffB = kOffALo + kKBlocks * kABlockBytes;          // 163840
constexpr int kOffEnc = kOffB + kBSlots * kBHalfBytes;            // 225280  float[2][192]: w0[64] w1[64] b[64]
constexpr int kOffBias = kOffEnc + 2 * 192 * 4;                   // 226816  float[4][256] of the CURRENT problem: b_in | b_r | b_z | b_hn
constexpr int kOffBar = kOffBias + 4 * 256 * 4;                   // 230912  mbarriers + TMEM base
constexpr int kSmemBytes = kOffBar + 256;                         // 231168 (+1024 alignment slack at launch)
static_assert(kSmemBytes + 1024 <= 232448, "shared memory budget of one sm_100 CTA");

struct TcState {
    __nv_bfloat16 *wimg;   // [2 problems][4 ct][5 kb][3 parts: bf16 hi | bf16 lo | fp16] x 24 KB swizzled images
    float *bias4;          // [2][4][256]: b_in | b_ir+b_hr | b_iz+b_hz | b_hn
    float *enc;            // [2][192]
    int num_sms;
    CUtensorMap wmap;      // the weight images as a [rows x 64] 16-bit tensor, box = 96 rows (one CTA's half of a chunk)
    void *encode_tiled;    // cuTensorMapEncodeTiled (resolved at run time): the resident-image path encodes two maps per forward
};

using namespace tc;

#ifdef EDGE_PROFILE
__device__ unsigned long long g_edge_prof[32];
#define PROF_DECL(n) long long n = 0
#define PROF_T0(t) const long long t = clock64()
#define PROF_ADD(n, t) n += clock64() - t
#define PROF_OUT(i, n) atomicAdd(&g_edge_prof[i], (unsigned long long)(n))
#define DBG(bit) (a.debug & (bit))
#else
#define DBG(bit) 0
#define PROF_DECL(n)
#define PROF_T0(t)
#define PROF_ADD(n, t)
#define PROF_OUT(i, n)
#endif

// ---------------------------------------------------------------------------------------------- weight packing
struct PackArgs {
    const float *w_ih[2], *w_hh[2], *b_ih[2], *b_hh[2], *enc_w[2], *enc_b[2];   // [0] spatial, [1] temporal
    __nv_bfloat16 *wimg;
    float *bias4, *enc;
};

__global__ void pack_edge_weights_kernel(const PackArgs a)
{
    const int total = 2 * kColTiles * kKBlocks * kBChunkRows * 64;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int t = idx;
        const int k = t & 63; t >>= 6;
        const int q = t % kBChunkRows; t /= kBChunkRows;
        const int kb = t % kKBlocks; t /= kKBlocks;
        const int ct = t % kColTiles; t /= kColTiles;
        const int p = t;
        const int j = ct * 64 + (q & 63);
        const int blk = q >> 6;                      // 0,1,2 within the 192-row chunk
        float w;
        if (kb == 0) {                               // rows [n_i ; r ; z] of W_ih
            const int gate_row = (blk == 0 ? 512 : blk == 1 ? 0 : 256) + j;
            w = a.w_ih[p][(size_t)gate_row * 64 + k];
        } else {                                     // rows [r ; z ; n_h] of W_hh
            const int gate_row = (blk == 0 ? 0 : blk == 1 ? 256 : 512) + j;
            w = a.w_hh[p][(size_t)gate_row * 256 + (kb - 1) * 64 + k];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const size_t chunk = ((size_t)(p * kColTiles + ct) * kKBlocks + kb) * 3;
        char *base = reinterpret_cast<char *>(a.wimg);
        *reinterpret_cast<__nv_bfloat16 *>(base + (chunk + 0) * kBChunkBytes + sw128_offset(q, k)) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(base + (chunk + 1) * kBChunkBytes + sw128_offset(q, k)) = lo;
        *reinterpret_cast<__half *>(base + (chunk + 2) * kBChunkBytes + sw128_offset(q, k)) = __float2half_rn(w);
    }
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * 256; idx += gridDim.x * blockDim.x) {
        const int p = idx >> 8, c = idx & 255;
        float *b = a.bias4 + (size_t)p * 4 * 256;
        b[0 * 256 + c] = a.b_ih[p][512 + c];
        b[1 * 256 + c] = a.b_ih[p][c] + a.b_hh[p][c];
        b[2 * 256 + c] = a.b_ih[p][256 + c] + a.b_hh[p][256 + c];
        b[3 * 256 + c] = a.b_hh[p][512 + c];
    }
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * 64; idx += gridDim.x * blockDim.x) {
        const int p = idx >> 6, k = idx & 63;
        float *e = a.enc + (size_t)p * 192;
        e[k] = a.enc_w[p][2 * k];
        e[64 + k] = a.enc_w[p][2 * k + 1];
        e[128 + k] = a.enc_b[p][k];
    }
}

// ---------------------------------------------------------------------------------------------- the kernel
struct EdgeTcArgs {
    const float *temporal_edges, *spatial_edges, *h_in, *masks;
    float *h_out;
    const __nv_bfloat16 *wimg;
    const float *bias4, *enc;
    int N, H;
    int tiles_spatial, tiles_temporal;   // 128-row tiles per problem
    int pairs_spatial, pairs_total;      // 256-row tile pairs (one per CTA pair and iteration); a pair never mixes problems
    int three_pass;
    int fp16;          // single pass with FP16 operands (images part 2) instead of BF16
    int debug;         // EDGE_PROFILE builds: what-if switches (timing only, results are wrong), 0 otherwise
    // training (one step of the masked GRU sequences of the PPO update, cn_dsrnn_edge_sequence_step): rows are addressed as
    // spatial row m -> in/out_off_s + m, temporal row m -> in/out_off_t + m (the [T*S | T*N] sequence layout) and the kernel
    // also writes what the backward needs: the gate values and the split-bf16 operands of the weight-gradient products
    int train;
    long long in_off_s, in_off_t, out_off_s, out_off_t;
    float *ws;                         // [rows, 4, 256] r | z | n | W_hn hm + b_hn
    __nv_bfloat16 *hm_hi, *hm_lo;      // [rows, 256] masked previous state
    __nv_bfloat16 *e_hi, *e_lo;        // [rows, 72] encoded input | 1 | 0 x 7
    // resident split-bf16 image of the hidden state, LOGICAL row order (spatial rows env * H + human, then N * H + env):
    // written by the epilogue when img_out_hi != NULL; read by TMA instead of the fp32 staging in the kImg instantiation
    __nv_bfloat16 *img_out_hi, *img_out_lo;
};

struct TileInfo { bool spatial; int p, row0, M; };

template <bool kTrain, bool kImg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
edge_gru_tc_kernel(const __grid_constant__ EdgeTcArgs a, const __grid_constant__ CUtensorMap wmap,
                   const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    float *s_enc = reinterpret_cast<float *>(smem + kOffEnc);
    float *s_bias = reinterpret_cast<float *>(smem + kOffBias);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kOffBar + 8 * kNumBars);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, stride = a.H + 1;
    const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the MMAs of the pair)
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    for (int i = threadIdx.x; i < 2 * 192; i += kThreads) s_enc[i] = a.enc[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kBSlots; ++i) {
            mbar_init(bar(kBarFull + i), 1);       // leader only: both halves of a weight chunk landed (24 KB of complete_tx)
            mbar_init(bar(kBarEmpty + i), 1);      // MMAs reading the slot retired (commit, both CTAs)
        }
        for (int i = 0; i < 3; ++i) {              // A image pieces: k-block 0 (encoded input) | k-blocks 1-2 | k-blocks 3-4
            mbar_init(bar(kBarAReady + i), 2 * kStageWarps);   // leader only
            mbar_init(bar(kBarAFree + i), 1);
        }
        mbar_init(bar(kBarHprev), kEpiWarps);     // this CTA's epilogue warps took h_prev of column tiles 2, 3 out of the A image
        mbar_init(bar(kBarALoaded), 1); mbar_init(bar(kBarALoaded + 1), 1);
        mbar_init(bar(kBarTmemFull), 1); mbar_init(bar(kBarTmemFull + 1), 1);
        mbar_init(bar(kBarTmemEmpty), 2 * kEpiWarps); mbar_init(bar(kBarTmemEmpty + 1), 2 * kEpiWarps);   // leader only
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    // accumulator columns of one buffer: n_i [0,64) | r [64,128) | z [128,192) | n_h [192,256).  Every MMA is N = 192: the
    // input block overwrites n_i|r|z, the four hidden blocks accumulate into r|z|n_h -- so n_h has to start from zero: the
    // epilogue clears it after reading (and here, once, before the first tile).
    if (warp < kEpiWarps) {
        const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 192u + (uint32_t)(warp >> 2) * 32u;
        tmem_st16_zero(t0); tmem_st16_zero(t0 + 16u); tmem_st16_zero(t0 + 256u); tmem_st16_zero(t0 + 256u + 16u);
        tmem_st_wait();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    // programmatic dependent launch: everything above (barriers, TMEM, the encoder table packed long before this forward)
    // may run beside the tail of the previous kernel; the hidden state, masks and edge inputs are read from here on
    pdl_wait();
    pdl_launch_dependents();

    auto tile_info = [&](int pair) {
        TileInfo t;
        t.spatial = pair < a.pairs_spatial;
        t.p = t.spatial ? 0 : 1;
        const int tile = 2 * (t.spatial ? pair : pair - a.pairs_spatial) + (int)rank;
        const bool valid = tile < (t.spatial ? a.tiles_spatial : a.tiles_temporal);
        t.row0 = tile * kRows;
        t.M = valid ? (t.spatial ? a.N * H : a.N) : 0;          // an odd tile count leaves the peer of the last pair idle
        return t;
    };
    auto mem_row_of = [&](const TileInfo &t, int m, int &env) {
        env = t.spatial ? m / H : m;
        return t.spatial ? env * stride + 1 + (m - env * H) : env * stride;
    };
    // sequence layout (training): input row (previous step / initial state) and output row (this step) of logical row m
    auto seq_in_row = [&](const TileInfo &t, int m) { return (size_t)((t.spatial ? a.in_off_s : a.in_off_t) + m); };
    auto seq_out_row = [&](const TileInfo &t, int m) { return (size_t)((t.spatial ? a.out_off_s : a.out_off_t) + m); };

    if (warp < kEpiWarps) {
        // =============================================================== epilogue warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
        const int tid = (warp & 3) * 32 + lane;   // row of the tile = TMEM lane this thread owns
        const int chalf = warp >> 2;              // warps 0-3: hidden units 0..31 of a column tile, warps 4-7: 32..63
        const uint32_t empty_remote = mapa_rank(bar(kBarTmemEmpty), 0);
        uint32_t ctg = 0;                         // column tiles consumed so far (selects the TMEM buffer / parity)
        int bias_p = -1;                          // problem whose biases are in shared memory
        PROF_DECL(p_waitfull); PROF_DECL(p_epi); PROF_DECL(p_hprev); PROF_DECL(p_total);
        PROF_T0(t_all);
        for (int pair = cluster_id; pair < a.pairs_total; pair += num_clusters) {
            const TileInfo t = tile_info(pair);
            const int m = t.row0 + tid;
            const bool ok = m < t.M;
            size_t mem_row = 0;
            if (ok) { int env; mem_row = kTrain ? seq_out_row(t, m) : (size_t)mem_row_of(t, m, env); }
            float *wsrow = kTrain ? a.ws + mem_row * 1024 : nullptr;
            const size_t img_row = (size_t)(t.spatial ? 0 : a.N * H) + (size_t)m;      // logical row of the resident image
            if (t.p != bias_p) {                  // at most twice per CTA: spatial pairs come first, then temporal ones
                asm volatile("bar.sync 1, 256;" ::: "memory");
                reinterpret_cast<float4 *>(s_bias)[threadIdx.x] = __ldg(reinterpret_cast<const float4 *>(a.bias4 + t.p * 4 * 256) + threadIdx.x);
                asm volatile("bar.sync 1, 256;" ::: "memory");
                bias_p = t.p;
            }
            float *orow = a.h_out + mem_row * 256;
            for (int ct = 0; ct < kColTiles; ++ct, ++ctg) {
                const uint32_t buf = ctg & 1u;
                const int cb = ct * 64 + chalf * 32;                 // first hidden unit handled by this warp
                PROF_T0(t_w);
                mbar_wait(bar(kBarTmemFull + buf), (ctg >> 1) & 1u);
                tc_fence_after();
                PROF_ADD(p_waitfull, t_w);
                // h_prev (already masked) of this row's 32 hidden units comes back out of the staged A image (bf16 hi + lo =
                // 16 mantissa bits; conflict-free 16-byte reads thanks to the 128B swizzle); the accumulators being complete implies
                // that the staging warps' writes of this tile are complete and visible.  The staging warps overwrite
                // k-blocks 0-2 only after the MMAs of column tile 3 passed k-block 2 -- those MMAs wait for the epilogue of
                // column tile 1, so the reads for tiles 0, 1 are long done -- and k-blocks 3-4 only after kBarHprev.
                PROF_T0(t_h);
                float hprev[32];
                {
                    const unsigned char *img = smem + (1 + (cb >> 6)) * kABlockBytes;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {                        // 4 chunks of 8 hidden units
                        const int off = sw128_offset(tid, (cb & 63) + q * 8);
                        const uint4 hi = *reinterpret_cast<const uint4 *>(img + kOffAHi + off);
                        uint4 lo = make_uint4(0u, 0u, 0u, 0u);
                        if (a.three_pass) lo = *reinterpret_cast<const uint4 *>(img + kOffALo + off);
                        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                        for (int w2 = 0; w2 < 4; ++w2) {
                            if (a.fp16) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hw[w2]));
                                hprev[q * 8 + 2 * w2] = f.x; hprev[q * 8 + 2 * w2 + 1] = f.y;
                            } else {
                                hprev[q * 8 + 2 * w2] = __uint_as_float(hw[w2] << 16) + __uint_as_float(lw[w2] << 16);
                                hprev[q * 8 + 2 * w2 + 1] = __uint_as_float(hw[w2] & 0xffff0000u) + __uint_as_float(lw[w2] & 0xffff0000u);
                            }
                        }
                    }
                }
                if (ct == kColTiles - 1) { __syncwarp(); if (lane == 0) mbar_arrive(bar(kBarHprev)); }
                PROF_ADD(p_hprev, t_h);
                PROF_T0(t_e);
                const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + buf * 256u;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c16 = chalf * 2 + cc;
                    float ni[16], rg[16], zg[16], nh[16];
                    tmem_ld16(t0 + 0 * 64 + c16 * 16, ni);
                    tmem_ld16(t0 + 1 * 64 + c16 * 16, rg);
                    tmem_ld16(t0 + 2 * 64 + c16 * 16, zg);
                    tmem_ld16(t0 + 3 * 64 + c16 * 16, nh);
                    tmem_ld_wait();
                    if (cc == 1) {
                        // everything this warp needs from the accumulator buffer is in registers now: clear n_h for the next column
                        // tile and hand the buffer back BEFORE the gate math and the stores of the second half, so the MMAs of the
                        // column tile after next can start that much earlier
                        tmem_st16_zero(t0 + 192u + (uint32_t)chalf * 32u);          // n_h starts the next column tile from zero
                        tmem_st16_zero(t0 + 192u + (uint32_t)chalf * 32u + 16u);
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote_cta(empty_remote + 8u * buf);
                    }
                    const int c0 = ct * 64 + c16 * 16;
                    if (ok && !DBG(4)) {
                        float o8[8], r8[8], z8[8], n8[8], h8[8];
                        uint32_t ih[8], il[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c = c0 + q * 4;
                            const float4 b_n = *reinterpret_cast<const float4 *>(s_bias + c), b_r = *reinterpret_cast<const float4 *>(s_bias + 256 + c),
                                         b_z = *reinterpret_cast<const float4 *>(s_bias + 512 + c), b_h = *reinterpret_cast<const float4 *>(s_bias + 768 + c);
                            const float bn[4] = {b_n.x, b_n.y, b_n.z, b_n.w}, br[4] = {b_r.x, b_r.y, b_r.z, b_r.w},
                                        bz[4] = {b_z.x, b_z.y, b_z.z, b_z.w}, bh[4] = {b_h.x, b_h.y, b_h.z, b_h.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int j = q * 4 + i;
                                const int o = (q & 1) * 4 + i;
                                if (kTrain) {
                                    h8[o] = nh[j] + bh[i];
                                    o8[o] = gru_blend_ws(rg[j] + br[i], zg[j] + bz[i], ni[j] + bn[i], h8[o], hprev[cc * 16 + j], r8[o], z8[o], n8[o]);
                                } else {
                                    o8[o] = gru_blend(rg[j] + br[i], zg[j] + bz[i], ni[j] + bn[i], nh[j] + bh[i], hprev[cc * 16 + j]);
                                }
                            }
                            // a row is 1 KB away from the next lane's row: every store instruction touches 32 lines, so the
                            // 32-byte form halves the LSU work of the epilogue
                            if (q & 1) {
                                st_global_v8(orow + c0 + (q - 1) * 4, o8);
                                if (!kTrain && a.img_out_hi) {          // the same values as split bf16, for the next step's TMA:
                                    const int w0 = (q >> 1) * 4;        // 16 columns = one 32-byte store per half (full sectors)
                                    split_bf16x2(o8[0], o8[1], ih[w0], il[w0]); split_bf16x2(o8[2], o8[3], ih[w0 + 1], il[w0 + 1]);
                                    split_bf16x2(o8[4], o8[5], ih[w0 + 2], il[w0 + 2]); split_bf16x2(o8[6], o8[7], ih[w0 + 3], il[w0 + 3]);
                                    if (q == 3) {
                                        const size_t io = img_row * 256 + (size_t)c0;
                                        st_global_v8_b32(a.img_out_hi + io, ih);
                                        st_global_v8_b32(a.img_out_lo + io, il);
                                    }
                                }
                                if (kTrain && !DBG(16)) {
                                    float *w8 = wsrow + c0 + (q - 1) * 4;
                                    st_global_v8(w8, r8); st_global_v8(w8 + 256, z8); st_global_v8(w8 + 512, n8); st_global_v8(w8 + 768, h8);
                                }
                            }
                        }
                    }
                }
                PROF_ADD(p_epi, t_e);
            }
        }
        PROF_ADD(p_total, t_all);
        if (threadIdx.x == 0) { PROF_OUT(1, p_waitfull); PROF_OUT(2, p_epi); PROF_OUT(3, p_hprev); PROF_OUT(4, p_total); PROF_OUT(5, 1); }
    } else if (warp < kEpiWarps + kStageWarps) {
        // =============================================================== staging warps: fp32 rows -> split-bf16 A image
        static_assert(kRegsStage == 96, "staging warps keep the launch allocation");
        const int sw = warp - kEpiWarps;          // this warp converts rows sw*16 .. sw*16+15 of the tile
        const uint32_t ready_remote = mapa_rank(bar(kBarAReady), 0);
        uint32_t it = 0;
        PROF_DECL(p_stage); PROF_DECL(p_wfree);
        for (int pair = cluster_id; pair < a.pairs_total; pair += num_clusters, ++it) {
            const TileInfo t = tile_info(pair);
            const float *enc = s_enc + t.p * 192;
            // lane b (and b+16) holds the bookkeeping of row b of this warp's 16 rows
            const int m_l = t.row0 + sw * 16 + (lane & 15);
            const bool ok_l = m_l < t.M;
            int ridx_l = 0;
            float mk_l = 0.f, x0_l = 0.f, x1_l = 0.f;
            long long orow_l = 0;                 // training: this row's index in the step's output arrays
            if (ok_l) {
                int env;
                ridx_l = mem_row_of(t, m_l, env);
                if (kTrain) { ridx_l = (int)seq_in_row(t, m_l); orow_l = (long long)seq_out_row(t, m_l); }
                mk_l = a.masks[env];
                const float *x = t.spatial ? a.spatial_edges + 2 * (size_t)m_l : a.temporal_edges + 2 * (size_t)m_l;
                x0_l = x[0]; x1_l = x[1];
            }
            if (DBG(128) && it > 0) {              // what-if (timing only): a perfect A staging -- hand every piece back at once
                for (int piece = 0; piece < 3; ++piece) {
                    mbar_wait(bar(kBarAFree + piece), (it - 1) & 1u);
                    if (piece == 2) mbar_wait(bar(kBarHprev), (it - 1) & 1u);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ready_remote + 8u * piece);
                }
                continue;
            }
            if (kImg) {
                // Resident-image path: k-blocks 1-4 arrive by TMA (warp kWarpAProducer) from the split-bf16 image the previous
                // forward's epilogue wrote; this warp only encodes k-block 0 and zeroes the rows of finished envs (mask 0: the
                // reference's h * mask) and the rows beyond the problem in shared memory before handing each piece on.
                PROF_T0(t_f0);
                if (it > 0) mbar_wait(bar(kBarAFree), (it - 1) & 1u);
                PROF_ADD(p_wfree, t_f0);
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const int r = sw * 16 + b;
                    const float x0 = __shfl_sync(0xffffffffu, x0_l, b), x1 = __shfl_sync(0xffffffffu, x1_l, b);
                    const bool okb = __shfl_sync(0xffffffffu, (int)ok_l, b) != 0;
                    float e[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int k = 2 * lane + i;
                        e[i] = okb ? fmaxf(fmaf(enc[64 + k], x1, enc[k] * x0) + enc[128 + k], 0.f) : 0.f;
                    }
                    uint32_t hi, lo;
                    split_bf16x2(e[0], e[1], hi, lo);
                    const int off = sw128_offset(r, 2 * lane);
                    *reinterpret_cast<uint32_t *>(smem + kOffAHi + off) = hi;
                    *reinterpret_cast<uint32_t *>(smem + kOffALo + off) = lo;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ready_remote);
                const unsigned dead = __ballot_sync(0xffffffffu, !ok_l || mk_l == 0.0f) & 0xffffu;      // lanes 0-15 <-> this warp's rows
#pragma unroll 1
                for (int p = 1; p <= 2; ++p) {
                    PROF_T0(t_f);
                    mbar_wait(bar(kBarALoaded + p - 1), it & 1u);
                    PROF_ADD(p_wfree, t_f);
                    for (unsigned todo = dead; todo; todo &= todo - 1u) {
                        const int r = sw * 16 + (__ffs(todo) - 1);
                        const int kb = 2 * p - 1 + (lane >> 4);
                        unsigned char *row = smem + ((lane >> 3) & 1 ? kOffALo : kOffAHi) + kb * kABlockBytes + (r >> 3) * 1024 + (r & 7) * 128;
                        *reinterpret_cast<uint4 *>(row + (lane & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ready_remote + 8u * p);
                }
                continue;
            }
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {   // hidden units [0,128) = k-blocks 1-2, then [128,256) = k-blocks 3-4
                float4 hv[16];
#pragma unroll
                for (int b = 0; b < 16; ++b) {        // 16 x 512 B in flight per warp before the first conversion
                    const int ridx = __shfl_sync(0xffffffffu, ridx_l, b);
                    const bool okb = __shfl_sync(0xffffffffu, (int)ok_l, b) != 0;
                    hv[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (okb && !DBG(8)) hv[b] = *reinterpret_cast<const float4 *>(a.h_in + (size_t)ridx * 256 + half * 128 + lane * 4);
                }
                // the loads above are in flight while the pair's MMAs on the previous tile finish reading the A image; its
                // pieces are released (and handed back) one by one so the next tile's first MMAs do not wait for all of it
                if (half == 0) {   // piece 0: encoded input block (k-block 0), no global data needed: lane owns k = 2*lane, 2*lane+1
                    PROF_T0(t_f0);
                    if (it > 0) mbar_wait(bar(kBarAFree), (it - 1) & 1u);
                    PROF_ADD(p_wfree, t_f0);
#pragma unroll
                    for (int b = 0; b < 16; ++b) {
                        const int r = sw * 16 + b;
                        const float x0 = __shfl_sync(0xffffffffu, x0_l, b), x1 = __shfl_sync(0xffffffffu, x1_l, b);
                        const bool okb = __shfl_sync(0xffffffffu, (int)ok_l, b) != 0;
                        float e[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int k = 2 * lane + i;
                            e[i] = okb ? fmaxf(fmaf(enc[64 + k], x1, enc[k] * x0) + enc[128 + k], 0.f) : 0.f;
                        }
                        uint32_t hi, lo;
                        if (a.fp16) { hi = pack_half2(e[0], e[1]); lo = 0u; } else split_bf16x2(e[0], e[1], hi, lo);
                        const int off = sw128_offset(r, 2 * lane);
                        *reinterpret_cast<uint32_t *>(smem + kOffAHi + off) = hi;
                        if (a.three_pass) *reinterpret_cast<uint32_t *>(smem + kOffALo + off) = lo;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ready_remote);
                }
                PROF_T0(t_f);
                if (it > 0) {
                    mbar_wait(bar(kBarAFree + 1 + half), (it - 1) & 1u);
                    if (half) mbar_wait(bar(kBarHprev), (it - 1) & 1u);
                }
                PROF_ADD(p_wfree, t_f);
                PROF_T0(t_s);
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const int r = sw * 16 + b;
                    const float mk = __shfl_sync(0xffffffffu, mk_l, b);
                    const int e0 = half * 128 + lane * 4;
                    float4 h4 = hv[b];
                    h4.x *= mk; h4.y *= mk; h4.z *= mk; h4.w *= mk;
                    const int off = (1 + (e0 >> 6)) * kABlockBytes + sw128_offset(r, e0 & 63);
                    uint2 hi, lo;
                    if (a.fp16) { hi.x = pack_half2(h4.x, h4.y); hi.y = pack_half2(h4.z, h4.w); lo.x = lo.y = 0u; }
                    else { split_bf16x2(h4.x, h4.y, hi.x, lo.x); split_bf16x2(h4.z, h4.w, hi.y, lo.y); }
                    *reinterpret_cast<uint2 *>(smem + kOffAHi + off) = hi;
                    if (a.three_pass) *reinterpret_cast<uint2 *>(smem + kOffALo + off) = lo;
                }
                fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor-core (async) proxy
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ready_remote + 8u * (1 + half));   // k-blocks 1-2, then 3-4, of this CTA's rows
                PROF_ADD(p_stage, t_s);
            }
            if (kTrain && !DBG(32)) {
                // Records of the backward, copied out of the staged A image AFTER all three pieces were handed to the MMA thread
                // (the image stays valid until this warp restages it; an arrive's release fence would otherwise wait for these
                // global stores): the split masked state (operand of dW_hh = dgh^T hm) and the split encoder output in rows of
                // 72 = 64 values, a constant 1 and 7 zeros -- the "ones column" that makes G^T [e | 1] also return the column
                // sums of G (the bias gradients).
#pragma unroll 2
                for (int b = 0; b < 16; ++b) {
                    const long long orow = __shfl_sync(0xffffffffu, orow_l, b);
                    if (!__shfl_sync(0xffffffffu, (int)ok_l, b)) continue;
                    const int r = sw * 16 + b;
                    const int offe = sw128_offset(r, 2 * lane);
                    reinterpret_cast<uint32_t *>(a.e_hi + orow * 72)[lane] = *reinterpret_cast<const uint32_t *>(smem + kOffAHi + offe);
                    reinterpret_cast<uint32_t *>(a.e_lo + orow * 72)[lane] = *reinterpret_cast<const uint32_t *>(smem + kOffALo + offe);
                    if (lane == 0) {
                        *reinterpret_cast<uint4 *>(a.e_hi + orow * 72 + 64) = make_uint4(0x00003F80u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4 *>(a.e_lo + orow * 72 + 64) = make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int e0 = half * 128 + lane * 4;
                        const int off = (1 + (e0 >> 6)) * kABlockBytes + sw128_offset(r, e0 & 63);
                        *reinterpret_cast<uint2 *>(a.hm_hi + orow * 256 + e0) = *reinterpret_cast<const uint2 *>(smem + kOffAHi + off);
                        *reinterpret_cast<uint2 *>(a.hm_lo + orow * 256 + e0) = *reinterpret_cast<const uint2 *>(smem + kOffALo + off);
                    }
                }
            }
            // pull the NEXT tile's hidden-state rows into L2 while this one is being multiplied (16 rows x 8 lines per warp)
            const int next = pair + num_clusters;
            if (next < a.pairs_total) {
                const TileInfo tn = tile_info(next);
                const int mn = tn.row0 + sw * 16 + (lane & 15);
                if (mn < tn.M) {
                    int env;
                    const size_t nrow = kTrain ? seq_in_row(tn, mn) : (size_t)mem_row_of(tn, mn, env);
                    const float *row = a.h_in + nrow * 256 + (lane >> 4) * 128;
#pragma unroll
                    for (int l = 0; l < 4; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + l * 32));
                }
            }
        }
        if (threadIdx.x == kEpiWarps * 32) { PROF_OUT(0, p_stage); PROF_OUT(6, p_wfree); }
    } else {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsMisc));
      if (warp == kWarpProducer) {
        // =============================================================== weight producer (one lane): this CTA's half of
        // every 192-row chunk = 96 weight rows = 12 KB, one TMA box; both CTAs' copies report to the LEADER's barrier
        // (cp.async.bulk.tensor .cta_group::2), so the MMA thread waits on one barrier per chunk
        if (lane == 0) {
            uint32_t slot = 0, empty_parity = 1;
            const int parts = a.three_pass ? 2 : 1;
            const uint32_t full_leader = mapa_rank(bar(kBarFull), 0);
            PROF_DECL(p_wait); PROF_DECL(p_total);
            PROF_T0(t_all);
            for (int pair = cluster_id; pair < a.pairs_total; pair += num_clusters) {
                const int p = pair < a.pairs_spatial ? 0 : 1;
                for (int ct = 0; ct < kColTiles; ++ct)
                    for (int kb = 0; kb < kKBlocks; ++kb)
                        for (int part = 0; part < parts; ++part) {
                            if (DBG(64) && kb > 0 && part == 1) continue;   // what-if (timing only): 2 pass-equivalents per hidden k-block
                            PROF_T0(tw);
                            mbar_wait(bar(kBarEmpty + slot), empty_parity);
                            PROF_ADD(p_wait, tw);
                            if (rank == 0) mbar_expect_tx(bar(kBarFull + slot), kBChunkBytes);
                            const int row = ((((p * kColTiles + ct) * kKBlocks + kb) * 3 + (a.fp16 ? 2 : part)) * kBChunkRows) + (int)rank * (kBChunkRows / 2);
                            tma_load_2d_pair(s_base + kOffB + slot * kBHalfBytes, &wmap, 0, row, full_leader + 8u * slot);
                            if (++slot == kBSlots) { slot = 0; empty_parity ^= 1u; }
                        }
            }
            PROF_ADD(p_total, t_all);
            PROF_OUT(16, p_wait); PROF_OUT(17, p_total);
        }
      } else if (warp == kWarpAProducer) {
        // =============================================================== resident-image path: TMA loads of the A image.
        // A piece (k-blocks 1-2 / 3-4: 2 x 2 boxes of 128 rows x 64 bf16) is fetched as soon as the MMAs of the previous tile
        // (and the epilogue's h_prev reads) are done with it.
        if (kImg && lane == 0) {
            uint32_t it = 0;
            for (int pair = cluster_id; pair < a.pairs_total; pair += num_clusters, ++it) {
                const TileInfo t = tile_info(pair);
                const int img_row0 = (t.spatial ? 0 : a.N * H) + t.row0;
                for (int p = 1; p <= 2; ++p) {
                    if (it > 0) {
                        mbar_wait(bar(kBarAFree + p), (it - 1) & 1u);
                        if (p == 2) mbar_wait(bar(kBarHprev), (it - 1) & 1u);
                    }
                    mbar_expect_tx(bar(kBarALoaded + p - 1), 4 * kABlockBytes);
                    for (int kb = 2 * p - 1; kb <= 2 * p; ++kb) {
                        tma_load_2d(s_base + kOffAHi + kb * kABlockBytes, &amap_hi, (kb - 1) * 64, img_row0, bar(kBarALoaded + p - 1));
                        tma_load_2d(s_base + kOffALo + kb * kABlockBytes, &amap_lo, (kb - 1) * 64, img_row0, bar(kBarALoaded + p - 1));
                    }
                }
            }
        }
      } else if (warp != kWarpMma) {
        // warp 19 only pads the last warpgroup
      } else if (rank != 0) {
        // the peer's lane of this warp has nothing to issue: cta_group::2 MMAs come from the leader alone
      } else {
        // =============================================================== MMA issuer (warp 17 of the leader CTA)
        // warp-uniform control flow; the elected lane issues (see elect_one_sync in tc_common.cuh)
        {
            const uint32_t lead = elect_one_sync();
            uint32_t ctg = 0, it = 0;
            const uint32_t id192 = a.fp16 ? idesc_f16_m256(192) : idesc_bf16_m256(192);
            // shared-memory descriptors: the high word is constant, the low word is (address >> 4) | LBO; stepping K by 16
            // elements (32 B) or to another k-block / ring slot only adds to the low word, so one MMA costs a few instructions.
            // The same offsets are valid in the peer CTA (identical shared-memory layout), which is what cta_group::2 reads.
            constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            auto make_desc = [](uint32_t lo) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi)); return d; };
            const uint32_t a_hi_desc_lo = ((s_base + kOffAHi) >> 4) | (1u << 16), a_lo_desc_lo = ((s_base + kOffALo) >> 4) | (1u << 16);
            const uint32_t b_desc_lo = ((s_base + kOffB) >> 4) | (1u << 16);
            PROF_DECL(p_wa); PROF_DECL(p_we); PROF_DECL(p_wb); PROF_DECL(p_total);
            PROF_T0(t_all);
            uint32_t slot = 0, full_parity = 0;       // ring position, advanced incrementally (no division in the loop)
            // four K=16 steps of one 64-wide k-block: D[dcol..+192) (+)= A(a_lo) * B(b_lo)^T on both CTAs of the pair
            auto mma_kblock = [&](uint32_t dcol, uint32_t a_lo, uint32_t b_lo, bool overwrite_first) {
                umma2_f16_if(lead, dcol, make_desc(a_lo), make_desc(b_lo), id192, overwrite_first ? 0u : 1u);
                umma2_f16_if(lead, dcol, make_desc(a_lo + 2), make_desc(b_lo + 2), id192, 1u);
                umma2_f16_if(lead, dcol, make_desc(a_lo + 4), make_desc(b_lo + 4), id192, 1u);
                umma2_f16_if(lead, dcol, make_desc(a_lo + 6), make_desc(b_lo + 6), id192, 1u);
            };
            auto next_chunk = [&](uint32_t &b_lo) {     // wait for the next weight chunk (both halves), return its descriptor word
                PROF_T0(twb);
                mbar_wait(bar(kBarFull + slot), full_parity);
                tc_fence_after();
                PROF_ADD(p_wb, twb);
                b_lo = b_desc_lo + slot * (kBHalfBytes >> 4);
            };
            PROF_DECL(p_commit);
            auto release_chunk = [&]() {
                PROF_T0(tc0);
                umma2_commit_pair_if(lead, bar(kBarEmpty + slot));
                PROF_ADD(p_commit, tc0);
                if (++slot == kBSlots) { slot = 0; full_parity ^= 1u; }
            };
            for (int pair = cluster_id; pair < a.pairs_total; pair += num_clusters, ++it) {
                for (int ct = 0; ct < kColTiles; ++ct, ++ctg) {
                    const uint32_t buf = ctg & 1u;
                    PROF_T0(twe);
                    mbar_wait_cluster(bar(kBarTmemEmpty + buf), ((ctg >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    PROF_ADD(p_we, twe);
                    const uint32_t d0 = tmem_base + buf * 256u;
#pragma unroll 1
                    for (int kb = 0; kb < kKBlocks; ++kb) {
                        if (ct == 0 && (kb == 0 || kb == 1 || kb == 3)) {      // both CTAs' staging warps finished k-block 0 / 1-2 / 3-4
                            PROF_T0(twa);
                            mbar_wait_cluster(bar(kBarAReady + (kb == 0 ? 0 : kb == 1 ? 1 : 2)), it & 1u);
                            tc_fence_after();
                            PROF_ADD(p_wa, twa);
                        }
                        const uint32_t dcol = d0 + (kb == 0 ? 0u : 64u);
                        const uint32_t a_hi = a_hi_desc_lo + kb * (kABlockBytes >> 4), a_lo = a_lo_desc_lo + kb * (kABlockBytes >> 4);
                        uint32_t b_lo;
                        next_chunk(b_lo);                                   // B_hi (bf16 hi, or the fp16 image)
                        if (!DBG(2)) {
                            mma_kblock(dcol, a_hi, b_lo, kb == 0);          // the input block overwrites n_i | r | z
                            if (a.three_pass) mma_kblock(dcol, a_lo, b_lo, false);
                        }
                        release_chunk();
                        if (a.three_pass && !(DBG(64) && kb > 0)) {
                            next_chunk(b_lo);                               // B_lo
                            if (!DBG(2)) mma_kblock(dcol, a_hi, b_lo, false);
                            release_chunk();
                        }
                        if (ct == kColTiles - 1 && kb == 0) umma2_commit_pair_if(lead, bar(kBarAFree));            // k-block 0 of A is free
                        if (ct == kColTiles - 1 && kb == 2) umma2_commit_pair_if(lead, bar(kBarAFree + 1));        // k-blocks 1-2 are free
                    }
                    umma2_commit_pair_if(lead, bar(kBarTmemFull + buf));     // accumulators of this column tile are complete in both CTAs
                }
                umma2_commit_pair_if(lead, bar(kBarAFree + 2));              // every MMA that reads this tile's A image has retired
            }
            PROF_ADD(p_total, t_all);
            if (lane == 0) { PROF_OUT(8, p_wa); PROF_OUT(9, p_we); PROF_OUT(10, p_wb); PROF_OUT(11, p_commit); PROF_OUT(12, p_total); }
        }
      }
    }

    tc_fence_before();
    cluster_sync_all();                     // the peer may still be signalling barriers in this CTA's shared memory
    if (warp == kWarpMma) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host side
#ifdef EDGE_PROFILE
// development builds only (make NET_FLAGS=-DEDGE_PROFILE): per-role cycle counters summed over CTAs, see tools/edge_profile.py
extern "C" int cn_debug_edge_profile(unsigned long long *out32, int reset)
{
    if (out32 && cudaMemcpyFromSymbol(out32, g_edge_prof, sizeof(g_edge_prof)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[32] = {}; if (cudaMemcpyToSymbol(g_edge_prof, z, sizeof(z)) != cudaSuccess) return 1; }
    return 0;
}
#endif
void dsrnn_tc_destroy(void *state);
const char *dsrnn_tc_repack(void *state, const CnDsrnnWeights *w, cudaStream_t stream);

const char *dsrnn_tc_create(const CnDsrnnWeights *w, cudaStream_t stream, void **state)
{
    *state = nullptr;
    TcState *st = new (std::nothrow) TcState();
    if (!st) return "out of host memory";
    const size_t img_bytes = (size_t)2 * kColTiles * kKBlocks * 3 * kBChunkBytes;
    if (cudaMalloc(&st->wimg, img_bytes) != cudaSuccess || cudaMalloc(&st->bias4, 2 * 4 * 256 * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&st->enc, 2 * 192 * sizeof(float)) != cudaSuccess) {
        delete st;
        return "cudaMalloc of the packed edge weights failed";
    }
    {   // the driver entry point is resolved at run time so that the library has no link-time dependency on libcuda
        typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            dsrnn_tc_destroy(st);
            return "cuTensorMapEncodeTiled is not available from this driver";
        }
        st->encode_tiled = fn;
        const cuuint64_t dims[2] = {64, (cuuint64_t)(img_bytes / 128)};      // 128-byte rows of the pre-swizzled images
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {64, kBChunkRows / 2}, estr[2] = {1, 1};
        if (reinterpret_cast<EncodeTiled>(fn)(&st->wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, st->wimg, dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            dsrnn_tc_destroy(st);
            return "cuTensorMapEncodeTiled failed for the edge weight images";
        }
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&st->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaFuncSetAttribute(edge_gru_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024) != cudaSuccess ||
        cudaFuncSetAttribute(edge_gru_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024) != cudaSuccess ||
        cudaFuncSetAttribute(edge_gru_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024) != cudaSuccess) {
        dsrnn_tc_destroy(st);
        return "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    }
    if (const char *msg = dsrnn_tc_repack(st, w, stream)) { dsrnn_tc_destroy(st); return msg; }
    *state = st;
    return nullptr;
}

// re-runs the pack kernel into the EXISTING images (same addresses: CUDA graphs that captured the forward stay valid,
// no allocation, no synchronisation); called after every optimiser step
const char *dsrnn_tc_repack(void *state, const CnDsrnnWeights *w, cudaStream_t stream)
{
    TcState *st = static_cast<TcState *>(state);
    if (!st) return "tensor-core edge stage was not initialised";
    PackArgs pa;
    pa.w_ih[0] = w->s_w_ih; pa.w_hh[0] = w->s_w_hh; pa.b_ih[0] = w->s_b_ih; pa.b_hh[0] = w->s_b_hh; pa.enc_w[0] = w->s_enc_w; pa.enc_b[0] = w->s_enc_b;
    pa.w_ih[1] = w->t_w_ih; pa.w_hh[1] = w->t_w_hh; pa.b_ih[1] = w->t_b_ih; pa.b_hh[1] = w->t_b_hh; pa.enc_w[1] = w->t_enc_w; pa.enc_b[1] = w->t_enc_b;
    pa.wimg = st->wimg; pa.bias4 = st->bias4; pa.enc = st->enc;
    pack_edge_weights_kernel<<<256, 256, 0, stream>>>(pa);
    return cudaGetLastError() == cudaSuccess ? nullptr : "pack_edge_weights_kernel launch failed";
}

void dsrnn_tc_destroy(void *state)
{
    TcState *st = static_cast<TcState *>(state);
    if (!st) return;
    cudaFree(st->wimg);
    cudaFree(st->bias4);
    cudaFree(st->enc);
    delete st;
}

// `img`: resident split-bf16 image of the hidden state in logical row order -- img[0], img[1] = hi / lo of THIS forward's input
// (valid only if the caller got them from the previous forward's img[2], img[3] for the same h_edge_in; masks must be 0 / 1),
// img[2], img[3] = where to write the output's image; NULL members switch the respective half off.
const char *dsrnn_tc_edge_forward(void *state, const CnDsrnnWeights *, int n_envs, int H, const CnDsrnnIO *io, int precision,
                                  cudaStream_t stream, int *launches, void *const img[4])
{
    TcState *st = static_cast<TcState *>(state);
    if (!st) return "tensor-core edge stage was not initialised";
    EdgeTcArgs a;
    a.temporal_edges = io->temporal_edges; a.spatial_edges = io->spatial_edges; a.h_in = io->h_edge_in; a.masks = io->masks;
    a.h_out = io->h_edge_out; a.wimg = st->wimg; a.bias4 = st->bias4; a.enc = st->enc;
    a.N = n_envs; a.H = H;
    a.tiles_spatial = (int)(((size_t)n_envs * H + kRows - 1) / kRows);
    a.tiles_temporal = (n_envs + kRows - 1) / kRows;
    a.pairs_spatial = (a.tiles_spatial + 1) / 2;
    a.pairs_total = a.pairs_spatial + (a.tiles_temporal + 1) / 2;
    a.three_pass = precision == CN_PREC_BF16X3 ? 1 : 0;
    a.fp16 = precision == CN_PREC_FP16 ? 1 : 0;
    a.debug = 0;
    a.train = 0;
    a.in_off_s = a.in_off_t = a.out_off_s = a.out_off_t = 0;
    a.ws = nullptr; a.hm_hi = a.hm_lo = a.e_hi = a.e_lo = nullptr;
#ifdef EDGE_PROFILE
    if (const char *dbg = getenv("CN_EDGE_DEBUG")) a.debug = atoi(dbg);
#endif
    const int max_pairs = st->num_sms / 2;                        // one CTA pair (cluster of 2) per TPC, persistent
    const int grid = 2 * (a.pairs_total < max_pairs ? a.pairs_total : max_pairs);
    const bool three = precision == CN_PREC_BF16X3;
    cudaError_t lerr = cudaSuccess;
    a.img_out_hi = three && img ? static_cast<__nv_bfloat16 *>(img[2]) : nullptr;
    a.img_out_lo = three && img ? static_cast<__nv_bfloat16 *>(img[3]) : nullptr;
    if (!a.img_out_lo) a.img_out_hi = nullptr;
    if (three && img && img[0] && img[1] && a.debug == 0) {
        typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        CUtensorMap mh, ml;
        const cuuint64_t dims[2] = {256, (cuuint64_t)n_envs * (cuuint64_t)(H + 1)};
        const cuuint64_t strides[1] = {512};
        const cuuint32_t box[2] = {64, (cuuint32_t)kRows}, estr[2] = {1, 1};
        for (int k = 0; k < 2; ++k)
            if (reinterpret_cast<EncodeTiled>(st->encode_tiled)(k ? &ml : &mh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, img[k], dims, strides, box, estr,
                                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return "cuTensorMapEncodeTiled failed for the hidden-state image";
        lerr = cn_launch(edge_gru_tc_kernel<false, true>, dim3(grid), dim3(kThreads), kSmemBytes + 1024, stream, CN_PDL_EDGE, a, st->wmap, mh, ml);
    } else {
        lerr = cn_launch(edge_gru_tc_kernel<false, false>, dim3(grid), dim3(kThreads), kSmemBytes + 1024, stream, CN_PDL_EDGE, a, st->wmap, st->wmap, st->wmap);
    }
    if (lerr != cudaSuccess) return cudaGetErrorString(lerr);
    ++*launches;
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

// One step t of the two masked edge-GRU sequences of the PPO update (both weight sets in one launch), bf16x3:
// h_out = GRUCell(ReLU(W_enc x_t + b), m_t * h_in) plus the records of the backward (see EdgeTcArgs).
const char *dsrnn_tc_edge_sequence_step(void *state, int n_envs, int H, const CnEdgeSeqStep *io, cudaStream_t stream)
{
    TcState *st = static_cast<TcState *>(state);
    if (!st) return "tensor-core edge stage was not initialised";
    EdgeTcArgs a;
    a.temporal_edges = io->temporal_edges; a.spatial_edges = io->spatial_edges; a.h_in = io->h_in; a.masks = io->masks;
    a.h_out = io->h_out; a.wimg = st->wimg; a.bias4 = st->bias4; a.enc = st->enc;
    a.N = n_envs; a.H = H;
    a.tiles_spatial = (int)(((size_t)n_envs * H + kRows - 1) / kRows);
    a.tiles_temporal = (n_envs + kRows - 1) / kRows;
    a.pairs_spatial = (a.tiles_spatial + 1) / 2;
    a.pairs_total = a.pairs_spatial + (a.tiles_temporal + 1) / 2;
    a.three_pass = 1; a.fp16 = 0; a.debug = 0;
#ifdef EDGE_PROFILE
    if (const char *dbg = getenv("CN_EDGE_DEBUG")) a.debug = atoi(dbg);
#endif
    a.train = 1;
    a.in_off_s = io->in_row_spatial; a.in_off_t = io->in_row_temporal;
    a.out_off_s = io->out_row_spatial; a.out_off_t = io->out_row_temporal;
    a.ws = io->ws;
    a.hm_hi = static_cast<__nv_bfloat16 *>(io->hm_hi); a.hm_lo = static_cast<__nv_bfloat16 *>(io->hm_lo);
    a.e_hi = static_cast<__nv_bfloat16 *>(io->e_hi); a.e_lo = static_cast<__nv_bfloat16 *>(io->e_lo);
    const int max_pairs = st->num_sms / 2;
    const int grid = 2 * (a.pairs_total < max_pairs ? a.pairs_total : max_pairs);
    a.img_out_hi = a.img_out_lo = nullptr;
    edge_gru_tc_kernel<true, false><<<grid, kThreads, kSmemBytes + 1024, stream>>>(a, st->wmap, st->wmap, st->wmap);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}
