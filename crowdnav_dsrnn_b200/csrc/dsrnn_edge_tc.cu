// dsrnn_edge_tc.cu -- K3 stage 1 on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
// The edge GRUs of the DS-RNN (HumanHumanEdgeRNN.forward, srnn_model.py:201-215; one GRU(64->256) step on
// N temporal + N*H spatial edge rows) are one tall GEMM per weight set:
//     [e | m*h]  (rows x 320)   x   [W_ih | W_hh]^T  (320 x 768)      e = ReLU(W_enc x + b)  (K = 2, computed in place)
// followed by the gate non-linearities.  One persistent CTA per SM walks 128-row tiles:
//   warps 0-7  stage the tile's A operand straight from the fp32 hidden state in HBM (masking, fp32 -> split bf16
//              hi/lo, 128B-swizzled K-major smem image), then act as the epilogue: tcgen05.ld the accumulators of a
//              64-hidden-unit column tile (n_i | r | z | n_h, 256 TMEM columns), apply sigmoid/tanh/blend and store h'.
//   warp 8     streams the pre-swizzled bf16 weight images (24 KB chunks) with cp.async.bulk into a 2-slot smem ring.
//   warp 9     owns TMEM (512 columns = 2 accumulator buffers) and issues tcgen05.mma.kind::f16 (M=128, N=192, K=16).
// Precision: CN_PREC_BF16X3 runs A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (fp32 accumulate in TMEM, ~2^-16 relative operand
// error); CN_PREC_BF16 runs the first pass only.
#include <new>
#include "dsrnn.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;                 // UMMA M
constexpr int kKBlocks = 5;                // 64-wide k-blocks: 1 encoded-input block + 4 hidden blocks
constexpr int kColTiles = 4;               // 64 hidden units per column tile
constexpr int kABlockBytes = kRows * 128;  // 128 rows x 64 bf16
constexpr int kBChunkRows = 192;                 // weight rows per (column tile, k-block, part): [n_i|r|z] or [r|z|n_h] x 64
constexpr int kBChunkBytes = kBChunkRows * 128;
constexpr int kBSlots = 2;
constexpr int kBarAReady = 2 * kBSlots, kBarTmemFull = kBarAReady + 1, kBarTmemEmpty = kBarTmemFull + 2;
constexpr int kThreads = 320;              // 8 staging/epilogue warps + weight producer + MMA issuer

constexpr int kOffAHi = 0;
constexpr int kOffALo = kOffAHi + kKBlocks * kABlockBytes;        //  81920
constexpr int kOffB = kOffALo + kKBlocks * kABlockBytes;          // 163840
constexpr int kOffBias = kOffB + kBSlots * kBChunkBytes;                // 212992  float[2][4][256]
constexpr int kOffEnc = kOffBias + 2 * 4 * 256 * 4;               // 221184  float[2][192]: w0[64] w1[64] b[64]
constexpr int kOffBar = kOffEnc + 2 * 192 * 4;                    // 222720  mbarriers
constexpr int kSmemBytes = kOffBar + 128;                         // 222848 (+1024 alignment slack at launch)

struct TcState {
    __nv_bfloat16 *wimg;   // [2 problems][4 ct][5 kb][3 parts: bf16 hi | bf16 lo | fp16] x 24 KB swizzled images
    float *bias4;          // [2][4][256]: b_in | b_ir+b_hr | b_iz+b_hz | b_hn
    float *enc;            // [2][192]
    int num_sms;
};

using namespace tc;

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
    int tiles_spatial, tiles_total;
    int three_pass;
    int fp16;          // single pass with FP16 operands (images part 2) instead of BF16
};

__global__ void __launch_bounds__(kThreads, 1) edge_gru_tc_kernel(const __grid_constant__ EdgeTcArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    float *s_bias = reinterpret_cast<float *>(smem + kOffBias);
    float *s_enc = reinterpret_cast<float *>(smem + kOffEnc);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kOffBar + 112);   // 13 mbarriers occupy the first 104 bytes
    // barrier map: 0,1 full_b | 2,3 empty_b | 4 a_ready | 5,6 tmem_full | 7,8 tmem_empty
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, stride = a.H + 1;

    for (int i = threadIdx.x; i < 2 * 4 * 256; i += kThreads) s_bias[i] = a.bias4[i];
    for (int i = threadIdx.x; i < 2 * 192; i += kThreads) s_enc[i] = a.enc[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * kBSlots; ++i) mbar_init(bar(i), 1);
        mbar_init(bar(kBarAReady), 256);
        mbar_init(bar(kBarTmemFull), 1); mbar_init(bar(kBarTmemFull + 1), 1);
        mbar_init(bar(kBarTmemEmpty), 256); mbar_init(bar(kBarTmemEmpty + 1), 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp < 8) {
        // =============================================================== staging + epilogue warps
        const int tid = (warp & 3) * 32 + lane;   // row of the tile (TMEM lane) this thread owns in the epilogue
        const int chalf = warp >> 2;              // warps 0-3: hidden units 0..31 of a column tile, warps 4-7: 32..63
        uint32_t ctg = 0;                         // column tiles consumed so far (selects the TMEM buffer / parity)

        struct TileInfo { bool spatial; int p, row0, M; };
        auto tile_info = [&](int tile) {
            TileInfo t;
            t.spatial = tile < a.tiles_spatial;
            t.p = t.spatial ? 0 : 1;
            t.row0 = (t.spatial ? tile : tile - a.tiles_spatial) * kRows;
            t.M = t.spatial ? a.N * H : a.N;
            return t;
        };
        auto mem_row_of = [&](const TileInfo &t, int m, int &env) {
            env = t.spatial ? m / H : m;
            return t.spatial ? (size_t)env * stride + 1 + (m - env * H) : (size_t)env * stride;
        };
        // pull the NEXT tile's hidden-state rows into L2 while this tile is being computed (128 rows x 8 lines / 256 threads)
        auto prefetch_tile = [&](const TileInfo &t) {
            const int r = threadIdx.x >> 1, m = t.row0 + r;
            if (m < t.M) {
                int env;
                const float *row = a.h_in + mem_row_of(t, m, env) * 256 + (threadIdx.x & 1) * 128;
#pragma unroll
                for (int l = 0; l < 4; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + l * 32));
            }
        };
        // stage A: warp w converts rows w*16 .. w*16+15, 8 rows per batch so 16 x 16 B loads per lane are in flight
        // before the first conversion (a warp reads one 1 KB row with two float4 loads per lane)
        auto stage_tile = [&](const TileInfo &t) {
            const float *enc = s_enc + t.p * 192;
#pragma unroll 1
            for (int rb = 0; rb < 16; rb += 8) {
                float4 hv[8][2];
                float xs[8][2], mks[8];
                bool oks[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int m = t.row0 + warp * 16 + rb + b;
                    oks[b] = m < t.M;
                    hv[b][0] = hv[b][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    xs[b][0] = xs[b][1] = mks[b] = 0.f;
                    if (oks[b]) {
                        int env;
                        const float *hrow = a.h_in + mem_row_of(t, m, env) * 256 + lane * 4;
                        hv[b][0] = *reinterpret_cast<const float4 *>(hrow);
                        hv[b][1] = *reinterpret_cast<const float4 *>(hrow + 128);
                        mks[b] = a.masks[env];
                        const float *x = t.spatial ? a.spatial_edges + 2 * (size_t)m : a.temporal_edges + 2 * (size_t)m;
                        xs[b][0] = x[0]; xs[b][1] = x[1];
                    }
                }
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int r = warp * 16 + rb + b;
                    {   // encoded input block (k-block 0): lane owns k = 2*lane, 2*lane+1
                        float e[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int k = 2 * lane + i;
                            e[i] = oks[b] ? fmaxf(fmaf(enc[64 + k], xs[b][1], enc[k] * xs[b][0]) + enc[128 + k], 0.f) : 0.f;
                        }
                        uint32_t hi, lo;
                        if (a.fp16) { hi = pack_half2(e[0], e[1]); lo = 0u; } else split_bf16x2(e[0], e[1], hi, lo);
                        const int off = sw128_offset(r, 2 * lane);
                        *reinterpret_cast<uint32_t *>(smem + kOffAHi + off) = hi;
                        if (a.three_pass) *reinterpret_cast<uint32_t *>(smem + kOffALo + off) = lo;
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {   // hidden blocks: elements half*128 + lane*4 .. +3
                        const int e0 = half * 128 + lane * 4;
                        float4 h4 = hv[b][half];
                        const float mk = mks[b];
                        h4.x *= mk; h4.y *= mk; h4.z *= mk; h4.w *= mk;
                        const int off = (1 + (e0 >> 6)) * kABlockBytes + sw128_offset(r, e0 & 63);
                        uint2 hi, lo;
                        if (a.fp16) { hi.x = pack_half2(h4.x, h4.y); hi.y = pack_half2(h4.z, h4.w); lo.x = lo.y = 0u; }
                        else { split_bf16x2(h4.x, h4.y, hi.x, lo.x); split_bf16x2(h4.z, h4.w, hi.y, lo.y); }
                        *reinterpret_cast<uint2 *>(smem + kOffAHi + off) = hi;
                        if (a.three_pass) *reinterpret_cast<uint2 *>(smem + kOffALo + off) = lo;
                    }
                }
            }
            fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            mbar_arrive(bar(kBarAReady));
        };

        int tile = blockIdx.x;
        if (tile < a.tiles_total) stage_tile(tile_info(tile));
        for (; tile < a.tiles_total; tile += gridDim.x) {
            const TileInfo t = tile_info(tile);
            const int next = tile + gridDim.x;
            const bool has_next = next < a.tiles_total;
            if (has_next) prefetch_tile(tile_info(next));
            // the epilogue reads h_prev back from the A image, whose rows were staged by OTHER warps: all staging stores
            // of this tile must have landed before any warp starts its first column tile
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- epilogue: thread tid owns row tid of the tile (TMEM lane tid)
            const int m = t.row0 + tid;
            const bool ok = m < t.M;
            size_t mem_row = 0;
            float mk = 0.f;
            if (ok) { int env; mem_row = mem_row_of(t, m, env); mk = a.masks[env]; }
            const float *bias = s_bias + t.p * 4 * 256;
            for (int ct = 0; ct < kColTiles; ++ct, ++ctg) {
                const uint32_t buf = ctg & 1u;
                // h_prev (already masked) of this row's 32 hidden units comes from the staged A image in shared memory
                // (bf16 hi + lo = 16 mantissa bits; conflict-free 16-byte reads thanks to the 128B swizzle) instead of
                // 8 row-strided global loads per thread
                float hprev[32];
                {
                    const int cb = ct * 64 + chalf * 32;                 // first hidden unit handled by this warp
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
                mbar_wait(bar(kBarTmemFull + buf), (ctg >> 1) & 1u);
                tc_fence_after();
                // the last column tile's accumulators are complete => every MMA that reads A has retired: restage A for
                // the next tile FIRST so its MMAs overlap this epilogue (the other TMEM buffer is already free)
                if (ct == kColTiles - 1 && has_next) {
                    asm volatile("bar.sync 1, 256;" ::: "memory");   // every warp has taken its h_prev out of the A image
                    stage_tile(tile_info(next));
                }
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
                    const int c0 = ct * 64 + c16 * 16;
                    if (ok) {
                        float4 *ho = reinterpret_cast<float4 *>(a.h_out + mem_row * 256 + c0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {

                            float o[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int c = c0 + q * 4 + i, j = q * 4 + i;
                                const float r = fast_sigmoid(rg[j] + bias[256 + c]);
                                const float z = fast_sigmoid(zg[j] + bias[512 + c]);
                                const float n = fast_tanh(ni[j] + bias[c] + r * (nh[j] + bias[768 + c]));
                                o[i] = (1.0f - z) * n + z * hprev[cc * 16 + q * 4 + i];
                            }
                            ho[q] = make_float4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(bar(kBarTmemEmpty + buf));
            }
        }
    } else if (warp == 8) {
        // =============================================================== weight producer (one lane)
        if (lane == 0) {
            uint32_t chunk = 0;
            const int parts = a.three_pass ? 2 : 1;
            for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x) {
                const int p = tile < a.tiles_spatial ? 0 : 1;
                const char *img = reinterpret_cast<const char *>(a.wimg) + (size_t)p * kColTiles * kKBlocks * 3 * kBChunkBytes;
                for (int ct = 0; ct < kColTiles; ++ct)
                    for (int kb = 0; kb < kKBlocks; ++kb)
                        for (int part = 0; part < parts; ++part, ++chunk) {
                            const uint32_t slot = chunk % kBSlots;
                            mbar_wait(bar(kBSlots + slot), ((chunk / kBSlots) & 1u) ^ 1u);
                            mbar_expect_tx(bar(slot), kBChunkBytes);
                            bulk_g2s(s_base + kOffB + slot * kBChunkBytes,
                                     img + ((size_t)(ct * kKBlocks + kb) * 3 + (a.fp16 ? 2 : part)) * kBChunkBytes, kBChunkBytes, bar(slot));
                        }
            }
        }
    } else {
        // =============================================================== MMA issuer (one lane)
        if (lane == 0) {
            uint32_t chunk = 0, ctg = 0, tile_iter = 0;
            const uint32_t id192 = a.fp16 ? idesc_f16(192) : idesc_bf16(192), id128 = a.fp16 ? idesc_f16(128) : idesc_bf16(128),
                           id64 = a.fp16 ? idesc_f16(64) : idesc_bf16(64);
            // shared-memory descriptors: the high word is constant, the low word is (address >> 4) | LBO; stepping K by 16
            // elements (32 B) or to another k-block / ring slot only adds to the low word, so one MMA costs a few instructions
            constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            auto make_desc = [](uint32_t lo) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi)); return d; };
            const uint32_t a_hi_desc_lo = ((s_base + kOffAHi) >> 4) | (1u << 16), a_lo_desc_lo = ((s_base + kOffALo) >> 4) | (1u << 16);
            const uint32_t b_desc_lo = ((s_base + kOffB) >> 4) | (1u << 16);
            for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x, ++tile_iter) {
                mbar_wait(bar(kBarAReady), tile_iter & 1u);
                tc_fence_after();
                for (int ct = 0; ct < kColTiles; ++ct, ++ctg) {
                    const uint32_t buf = ctg & 1u;
                    mbar_wait(bar(kBarTmemEmpty + buf), ((ctg >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d0 = tmem_base + buf * 256u;
                    for (int kb = 0; kb < kKBlocks; ++kb) {
                        const uint32_t dcol = d0 + (kb == 0 ? 0u : 64u);
                        const int parts = a.three_pass ? 2 : 1;
                        for (int part = 0; part < parts; ++part, ++chunk) {   // part 0: B_hi (passes A_hi, A_lo), part 1: B_lo (pass A_hi)
                            const uint32_t slot = chunk % kBSlots;
                            mbar_wait(bar(slot), (chunk / kBSlots) & 1u);
                            tc_fence_after();
                            const uint32_t b_lo = b_desc_lo + slot * (kBChunkBytes >> 4);
                            const int passes = (part == 0 && a.three_pass) ? 2 : 1;
                            for (int ps = 0; ps < passes; ++ps) {
                                const uint32_t a_lo = (ps == 0 ? a_hi_desc_lo : a_lo_desc_lo) + kb * (kABlockBytes >> 4);
#pragma unroll
                                for (int k16 = 0; k16 < 4; ++k16) {
                                    const uint64_t ad = make_desc(a_lo + 2 * k16), bd = make_desc(b_lo + 2 * k16);
                                    const bool first = part == 0 && ps == 0 && k16 == 0;
                                    if (first && kb == 0) umma_bf16(dcol, ad, bd, id192, 0u);           // overwrite n_i | r | z
                                    else if (first && kb == 1) {                                         // r | z accumulate, n_h starts
                                        umma_bf16(dcol, ad, bd, id128, 1u);
                                        umma_bf16(dcol + 128u, ad, make_desc(b_lo + ((128 * 128) >> 4) + 2 * k16), id64, 0u);
                                    } else umma_bf16(dcol, ad, bd, id192, 1u);
                                }
                            }
                            umma_commit(bar(kBSlots + slot));
                        }
                    }
                    umma_commit(bar(kBarTmemFull + buf));       // accumulators of this column tile are complete
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host side
void dsrnn_tc_destroy(void *state);

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
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&st->num_sms, cudaDevAttrMultiProcessorCount, dev);
    PackArgs pa;
    pa.w_ih[0] = w->s_w_ih; pa.w_hh[0] = w->s_w_hh; pa.b_ih[0] = w->s_b_ih; pa.b_hh[0] = w->s_b_hh; pa.enc_w[0] = w->s_enc_w; pa.enc_b[0] = w->s_enc_b;
    pa.w_ih[1] = w->t_w_ih; pa.w_hh[1] = w->t_w_hh; pa.b_ih[1] = w->t_b_ih; pa.b_hh[1] = w->t_b_hh; pa.enc_w[1] = w->t_enc_w; pa.enc_b[1] = w->t_enc_b;
    pa.wimg = st->wimg; pa.bias4 = st->bias4; pa.enc = st->enc;
    pack_edge_weights_kernel<<<256, 256, 0, stream>>>(pa);
    if (cudaGetLastError() != cudaSuccess) { dsrnn_tc_destroy(st); return "pack_edge_weights_kernel launch failed"; }
    if (cudaFuncSetAttribute(edge_gru_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024) != cudaSuccess) {
        dsrnn_tc_destroy(st);
        return "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    }
    *state = st;
    return nullptr;
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

const char *dsrnn_tc_edge_forward(void *state, const CnDsrnnWeights *, int n_envs, int H, const CnDsrnnIO *io, int precision,
                                  cudaStream_t stream, int *launches)
{
    TcState *st = static_cast<TcState *>(state);
    if (!st) return "tensor-core edge stage was not initialised";
    EdgeTcArgs a;
    a.temporal_edges = io->temporal_edges; a.spatial_edges = io->spatial_edges; a.h_in = io->h_edge_in; a.masks = io->masks;
    a.h_out = io->h_edge_out; a.wimg = st->wimg; a.bias4 = st->bias4; a.enc = st->enc;
    a.N = n_envs; a.H = H;
    a.tiles_spatial = (int)(((size_t)n_envs * H + kRows - 1) / kRows);
    a.tiles_total = a.tiles_spatial + (n_envs + kRows - 1) / kRows;
    a.three_pass = precision == CN_PREC_BF16X3 ? 1 : 0;
    a.fp16 = precision == CN_PREC_FP16 ? 1 : 0;
    const int grid = a.tiles_total < st->num_sms ? a.tiles_total : st->num_sms;
    edge_gru_tc_kernel<<<grid, kThreads, kSmemBytes + 1024, stream>>>(a);
    ++*launches;
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}
