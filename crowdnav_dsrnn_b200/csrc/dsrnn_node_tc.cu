// dsrnn_node_tc.cu -- K3 stages 3 + 4 (node RNN and actor / critic heads) as ONE tcgen05 kernel, sm_100a only.
//
// After the attention kernel every env owns cat = [o_t | c] (512 values).  What is left of SRNN.forward
// (srnn_model.py:149-173 HumanNodeRNN, :378-395/:487-495 actor / critic, distributions.py:85-94 fc_mean) is a chain of
// seven small dense layers per env.  As separate launches each of them is one 128-row tile per CTA with no overlap at
// all (35-65 us per launch, ten launches); here one CTA walks the whole chain for its 128 envs with the activations
// never leaving the SM:
//
//   epoch 0,1  emb  = ReLU(W_na cat + b)                      K = 512 (cat streamed from HBM in two halves), N = 64
//   epoch 2    node GRU(128 -> 128): A = [enc | emb | m*h], accumulators n_i | r | z | n_h (4 x 128 TMEM columns)
//   epoch 3    y    = W_out h'                                 K = 128, N = 256
//   epoch 4    [a1 | c1] = tanh([W_a0 ; W_c0] y + b)            K = 256, N = 512 (all of TMEM)
//   epoch 5    feat = tanh(W_a2 a1 + b)  -> action mean        K = 256, N = 256
//   epoch 6    c2   = tanh(W_c2 c1 + b)  -> value              K = 256, N = 256
//
// Roles: warps 0-7 stage / convert the A operand (fp32 -> split bf16 hi|lo, 128B-swizzled K-major images, four 64-wide
// k-blocks) and run every epilogue (tcgen05.ld -> bias -> activation -> next A image or the outputs); warp 8 streams the
// pre-swizzled weight chunks (<= 256 rows x 64 K, 32 KB) through a two-slot ring with cp.async.bulk; warp 9 issues the
// MMAs (M = 128, N = chunk rows).  The epochs of one tile are strictly sequential (each needs the previous result), the
// weight stream is not: it keeps flowing while the epilogues run, and at 1.7 MB per tile it is what bounds the kernel.
#include <new>
#include "dsrnn.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kRows = 128;
constexpr int kABlockBytes = kRows * 128;      // one 64-wide k-block of one part (hi or lo)
constexpr int kAKb = 4;                        // k-block slots of the A region
constexpr int kBSlotBytes = 256 * 128;         // a weight chunk: up to 256 rows x 64 K of one part
constexpr int kBSlots = 2;
constexpr int kThreads = 320;
constexpr int kNumChunks = 34;

// float table staged in shared memory (biases, head weights, the tiny robot encoder)
constexpr int kCEmbB = 0, kCGruB = kCEmbB + 64, kCOutB = kCGruB + 512, kCAc0B = kCOutB + 256, kCA2B = kCAc0B + 512,
              kCC2B = kCA2B + 256, kCWv = kCC2B + 256, kCWm = kCWv + 256, kCRw = kCWm + 512, kCRb = kCRw + 21, kCEw = kCRb + 3,
              kCEb = kCEw + 192, kCBv = kCEb + 64, kCBm = kCBv + 1, kCTotal = ((kCBm + 2 + 3) / 4) * 4;

constexpr int kOffAHi = 0;
constexpr int kOffALo = kOffAHi + kAKb * kABlockBytes;       //  65536
constexpr int kOffB = kOffALo + kAKb * kABlockBytes;         // 131072
constexpr int kOffConst = kOffB + kBSlots * kBSlotBytes;     // 196608
constexpr int kOffPart = kOffConst + kCTotal * 4;            // partial head dot products [2][128][3]
constexpr int kOffBar = kOffPart + 2 * 128 * 3 * 4;
constexpr int kSmemBytes = kOffBar + 128;
static_assert(kSmemBytes + 1024 <= 232448, "shared memory budget of one sm_100 CTA");
// barriers: 0,1 full_b | 2,3 empty_b | 4 a_ready (256 arrivals) | 5 epoch_done (commit)
constexpr int kBarFull = 0, kBarEmpty = kBSlots, kBarAReady = 2 * kBSlots, kBarEpochDone = kBarAReady + 1;

struct NodeChunk { unsigned char epoch, a_kb, overwrite, pad; unsigned short n, dcol; };
__constant__ NodeChunk c_chunks[kNumChunks];

struct NodeTcState {
    __nv_bfloat16 *wimg;   // [34 chunks][3 parts: bf16 hi | bf16 lo | fp16] x 32 KB swizzled images
    float *consts;         // kCTotal floats
    int num_sms;
};

// ---------------------------------------------------------------------------------------------- weight packing
struct PackSeg { const float *w; int ld, src_row, k0, dst_row, count; };
struct PackChunk { PackSeg seg[2]; int nseg; };
struct PackTable { PackChunk c[kNumChunks]; };

static_assert(sizeof(PackTable) <= 3584, "the pack table travels as a kernel parameter");

__global__ void pack_node_weights_kernel(const __grid_constant__ PackTable table, __nv_bfloat16 *__restrict__ wimg)
{
    const int total = kNumChunks * 256 * 64;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int k = idx & 63, row = (idx >> 6) & 255, ch = idx >> 14;
        const PackChunk &pc = table.c[ch];
        float w = 0.0f;
        for (int s = 0; s < pc.nseg; ++s) {
            const PackSeg &g = pc.seg[s];
            if (row >= g.dst_row && row < g.dst_row + g.count) w = g.w[(size_t)(g.src_row + row - g.dst_row) * g.ld + g.k0 + k];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        char *base = reinterpret_cast<char *>(wimg) + (size_t)ch * 3 * kBSlotBytes + sw128_offset(row, k);
        *reinterpret_cast<__nv_bfloat16 *>(base) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(base + kBSlotBytes) = lo;
        *reinterpret_cast<__half *>(base + 2 * kBSlotBytes) = __float2half_rn(w);
    }
}

__global__ void pack_node_consts_kernel(const CnDsrnnWeights w, float *__restrict__ c)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kCTotal; i += gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (i < kCGruB) v = w.n_att_b[i];
        else if (i < kCOutB) {                       // n_i: b_in | r: b_ir + b_hr | z: b_iz + b_hz | n_h: b_hn   (torch gate order r, z, n)
            const int g = (i - kCGruB) >> 7, j = (i - kCGruB) & 127;
            v = g == 0 ? w.n_b_ih[256 + j] : g == 1 ? w.n_b_ih[j] + w.n_b_hh[j] : g == 2 ? w.n_b_ih[128 + j] + w.n_b_hh[128 + j] : w.n_b_hh[256 + j];
        } else if (i < kCAc0B) v = w.n_out_b[i - kCOutB];
        else if (i < kCA2B) v = (i - kCAc0B) < 256 ? w.actor0_b[i - kCAc0B] : w.critic0_b[i - kCAc0B - 256];
        else if (i < kCC2B) v = w.actor2_b[i - kCA2B];
        else if (i < kCWv) v = w.critic2_b[i - kCC2B];
        else if (i < kCWm) v = w.critic_lin_w[i - kCWv];
        else if (i < kCRw) v = w.mean_w[i - kCWm];
        else if (i < kCRb) v = w.robot_w[i - kCRw];
        else if (i < kCEw) v = w.robot_b[i - kCRb];
        else if (i < kCEb) v = w.n_enc_w[i - kCEw];
        else if (i < kCBv) v = w.n_enc_b[i - kCEb];
        else if (i < kCBm) v = w.critic_lin_b[0];
        else if (i < kCBm + 2) v = w.mean_b[i - kCBm];
        c[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------- the kernel
struct NodeTcArgs {
    const float *cat, *robot_node, *h_node_in, *masks;
    float *h_node_out, *value, *action_mean, *feat;
    const __nv_bfloat16 *wimg;
    const float *consts;
    int N, tiles;
    int three_pass, fp16;
};

__global__ void __launch_bounds__(kThreads, 1) node_heads_tc_kernel(const __grid_constant__ NodeTcArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    float *s_c = reinterpret_cast<float *>(smem + kOffConst);
    float *s_part = reinterpret_cast<float *>(smem + kOffPart);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kOffBar + 64);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < kCTotal; i += kThreads) s_c[i] = a.consts[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * kBSlots; ++i) mbar_init(bar(i), 1);
        mbar_init(bar(kBarAReady), 256);
        mbar_init(bar(kBarEpochDone), 1);
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
    const int parts = a.three_pass ? 2 : 1;

    if (warp < 8) {
        // =============================================================== staging + epilogue warps
        const int tid = (warp & 3) * 32 + lane;   // row of the tile = TMEM lane this thread owns in the epilogues
        const int chalf = warp >> 2;              // interleaved 16-column chunks: warps 0-3 take the even ones, 4-7 the odd ones
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t done_count = 0;                  // epochs of this CTA whose MMAs have been waited for

        // write 16 consecutive K elements (k0 multiple of 16) of row `r` into the A image (split bf16 hi | lo, or fp16)
        auto put_a16 = [&](int r, int kcol, const float *v) {
            const int kb = kcol >> 6, k = kcol & 63;
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
                uint4 hi, lo;
                const float *p = v + h8 * 8;
                if (a.fp16) {
                    hi = make_uint4(pack_half2(p[0], p[1]), pack_half2(p[2], p[3]), pack_half2(p[4], p[5]), pack_half2(p[6], p[7]));
                    lo = make_uint4(0u, 0u, 0u, 0u);
                } else {
                    split_bf16x2(p[0], p[1], hi.x, lo.x); split_bf16x2(p[2], p[3], hi.y, lo.y);
                    split_bf16x2(p[4], p[5], hi.z, lo.z); split_bf16x2(p[6], p[7], hi.w, lo.w);
                }
                const int off = kb * kABlockBytes + sw128_offset(r, k + h8 * 8);
                *reinterpret_cast<uint4 *>(smem + kOffAHi + off) = hi;
                if (a.three_pass) *reinterpret_cast<uint4 *>(smem + kOffALo + off) = lo;
            }
        };
        // coalesced staging of `nkb` k-blocks from a row-major fp32 matrix: warp w converts rows w*16 .. w*16+15,
        // a lane owns 4 consecutive columns of a 128-column (two k-block) stripe
        auto stage_global = [&](const float *src, int ld, int col0, int nkb, int dst_kb0, int row0, bool use_mask) {
            for (int stripe = 0; stripe < nkb / 2; ++stripe) {
#pragma unroll 1
                for (int rb = 0; rb < 16; rb += 8) {
                    float4 hv[8];
                    float mk[8];
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const int e = row0 + warp * 16 + rb + b;
                        hv[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                        mk[b] = 1.0f;
                        if (e < a.N) {
                            hv[b] = *reinterpret_cast<const float4 *>(src + (size_t)e * ld + col0 + stripe * 128 + lane * 4);
                            if (use_mask) mk[b] = a.masks[e];
                        }
                    }
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const int r = warp * 16 + rb + b;
                        float4 h4 = hv[b];
                        h4.x *= mk[b]; h4.y *= mk[b]; h4.z *= mk[b]; h4.w *= mk[b];
                        const int e0 = stripe * 128 + lane * 4;
                        const int off = (dst_kb0 + (e0 >> 6)) * kABlockBytes + sw128_offset(r, e0 & 63);
                        uint2 hi, lo;
                        if (a.fp16) { hi.x = pack_half2(h4.x, h4.y); hi.y = pack_half2(h4.z, h4.w); lo.x = lo.y = 0u; }
                        else { split_bf16x2(h4.x, h4.y, hi.x, lo.x); split_bf16x2(h4.z, h4.w, hi.y, lo.y); }
                        *reinterpret_cast<uint2 *>(smem + kOffAHi + off) = hi;
                        if (a.three_pass) *reinterpret_cast<uint2 *>(smem + kOffALo + off) = lo;
                    }
                }
            }
        };
        auto a_ready = [&]() {       // this thread's part of the next A image is written and its TMEM reads are done
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar(kBarAReady));
        };
        auto wait_epoch = [&]() {    // all MMAs of the current epoch have retired: accumulators complete, A image free
            mbar_wait(bar(kBarEpochDone), done_count & 1u);
            ++done_count;
            tc_fence_after();
        };
        // act(acc + bias) of this thread's column chunks of a [128 x ncols] accumulator -> next A image at K = column
        auto epilogue_to_a = [&](int tcol0, int ncols, const float *bias, int act) {
            for (int c0 = chalf * 16; c0 < ncols; c0 += 32) {
                float v[16];
                tmem_ld16(t_lane + (uint32_t)(tcol0 + c0), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x = v[j] + bias[c0 + j];
                    if (act == 1) x = fmaxf(x, 0.0f); else if (act == 2) x = fast_tanh(x);
                    v[j] = x;
                }
                put_a16(tid, c0, v);
            }
        };

        // the prologue above, the constant table and the weight stream of warp 8 (packed long before this forward) overlap the
        // tail of the attention kernel; its output is read from here on
        pdl_wait();
        for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
            const int row0 = tile * kRows;
            const int e = row0 + tid;
            const bool ok = e < a.N;
            // ---- epoch 0 / 1: cat[:, 0:256], cat[:, 256:512] -> emb accumulator (64 columns)
            stage_global(a.cat, 512, 0, 4, 0, row0, false);
            a_ready();
            wait_epoch();
            stage_global(a.cat, 512, 256, 4, 0, row0, false);
            a_ready();
            wait_epoch();
            // ---- epoch 2 operand: x = [enc | emb] in k-blocks 0, 1 and the masked h_node in k-blocks 2, 3
            {
                // emb = ReLU(acc + b): 64 columns -> K 64..127 (k-block 1)
                for (int c0 = chalf * 16; c0 < 64; c0 += 32) {
                    float v[16];
                    tmem_ld16(t_lane + (uint32_t)c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + s_c[kCEmbB + c0 + j], 0.0f);
                    put_a16(tid, 64 + c0, v);
                }
                // enc = ReLU(W_ne (W_r robot_node + b_r) + b_ne): 64 columns -> k-block 0 (robot_linear 7->3, encoder_linear 3->64)
                float r3[3] = {0.f, 0.f, 0.f};
                if (ok) {
                    const float *rn = a.robot_node + (size_t)e * 7;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        float s = 0.0f;
#pragma unroll
                        for (int i = 0; i < 7; ++i) s = fmaf(s_c[kCRw + j * 7 + i], rn[i], s);
                        r3[j] = s + s_c[kCRb + j];
                    }
                }
                for (int c0 = chalf * 16; c0 < 64; c0 += 32) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = c0 + j;
                        float s = 0.0f;
#pragma unroll
                        for (int q = 0; q < 3; ++q) s = fmaf(s_c[kCEw + k * 3 + q], r3[q], s);
                        v[j] = ok ? fmaxf(s + s_c[kCEb + k], 0.0f) : 0.0f;
                    }
                    put_a16(tid, c0, v);
                }
                stage_global(a.h_node_in, 128, 0, 2, 2, row0, true);
            }
            a_ready();
            wait_epoch();
            // ---- node GRU gates: accumulators n_i [0,128) | r [128,256) | z [256,384) | n_h [384,512); h' -> HBM and k-blocks 0, 1
            for (int c0 = chalf * 16; c0 < 128; c0 += 32) {
                float ni[16], rg[16], zg[16], nh[16], hp[16];
                tmem_ld16(t_lane + (uint32_t)c0, ni);
                tmem_ld16(t_lane + (uint32_t)(128 + c0), rg);
                tmem_ld16(t_lane + (uint32_t)(256 + c0), zg);
                tmem_ld16(t_lane + (uint32_t)(384 + c0), nh);
                {   // masked h_prev back out of the A image (k-blocks 2, 3)
                    const unsigned char *img = smem + (2 + (c0 >> 6)) * kABlockBytes;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int off = sw128_offset(tid, (c0 & 63) + q * 8);
                        const uint4 hi = *reinterpret_cast<const uint4 *>(img + kOffAHi + off);
                        uint4 lo = make_uint4(0u, 0u, 0u, 0u);
                        if (a.three_pass) lo = *reinterpret_cast<const uint4 *>(img + kOffALo + off);
                        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                        for (int w2 = 0; w2 < 4; ++w2) {
                            if (a.fp16) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hw[w2]));
                                hp[q * 8 + 2 * w2] = f.x; hp[q * 8 + 2 * w2 + 1] = f.y;
                            } else {
                                hp[q * 8 + 2 * w2] = __uint_as_float(hw[w2] << 16) + __uint_as_float(lw[w2] << 16);
                                hp[q * 8 + 2 * w2 + 1] = __uint_as_float(hw[w2] & 0xffff0000u) + __uint_as_float(lw[w2] & 0xffff0000u);
                            }
                        }
                    }
                }
                tmem_ld_wait();
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = c0 + j;
                    o[j] = gru_blend(rg[j] + s_c[kCGruB + 128 + c], zg[j] + s_c[kCGruB + 256 + c], ni[j] + s_c[kCGruB + c],
                                     nh[j] + s_c[kCGruB + 384 + c], hp[j]);
                }
                if (ok) {
                    float4 *dst = reinterpret_cast<float4 *>(a.h_node_out + (size_t)e * 128 + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                }
                put_a16(tid, c0, o);
            }
            a_ready();
            wait_epoch();
            // ---- y = acc + b (256 columns) -> k-blocks 0..3
            epilogue_to_a(0, 256, s_c + kCOutB, 0);
            a_ready();
            wait_epoch();
            // ---- a1 = tanh(actor.0): accumulator columns [0,256) -> A; c1 stays in columns [256,512) for later
            epilogue_to_a(0, 256, s_c + kCAc0B, 2);
            a_ready();
            wait_epoch();
            // ---- feat = tanh(actor.2) -> fc_mean partial dot products; then c1 = tanh(critic.0) -> A
            float m0 = 0.0f, m1 = 0.0f, vsum = 0.0f;
            for (int c0 = chalf * 16; c0 < 256; c0 += 32) {
                float v[16];
                tmem_ld16(t_lane + (uint32_t)c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = fast_tanh(v[j] + s_c[kCA2B + c0 + j]);
                    m0 = fmaf(v[j], s_c[kCWm + c0 + j], m0);
                    m1 = fmaf(v[j], s_c[kCWm + 256 + c0 + j], m1);
                }
                if (a.feat && ok) {
                    float4 *dst = reinterpret_cast<float4 *>(a.feat + (size_t)e * 256 + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                }
            }
            epilogue_to_a(256, 256, s_c + kCAc0B + 256, 2);
            a_ready();
            wait_epoch();
            // ---- c2 = tanh(critic.2) -> critic_linear partial dot product
            for (int c0 = chalf * 16; c0 < 256; c0 += 32) {
                float v[16];
                tmem_ld16(t_lane + (uint32_t)c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) vsum = fmaf(fast_tanh(v[j] + s_c[kCC2B + c0 + j]), s_c[kCWv + c0 + j], vsum);
            }
            tc_fence_before();
            // the two column halves of a row live in different warps: combine through shared memory
            s_part[(chalf * 128 + tid) * 3 + 0] = vsum; s_part[(chalf * 128 + tid) * 3 + 1] = m0; s_part[(chalf * 128 + tid) * 3 + 2] = m1;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (chalf == 0 && ok) {
                const float *p0 = s_part + tid * 3, *p1 = s_part + (128 + tid) * 3;
                a.value[e] = p0[0] + p1[0] + s_c[kCBv];
                a.action_mean[2 * (size_t)e] = p0[1] + p1[1] + s_c[kCBm];
                a.action_mean[2 * (size_t)e + 1] = p0[2] + p1[2] + s_c[kCBm + 1];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");      // s_part is reused by the next tile
        }
    } else if (warp == 8) {
        // =============================================================== weight producer (one lane)
        if (lane == 0) {
            uint32_t slot = 0, empty_parity = 1;
            for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x)
                for (int c = 0; c < kNumChunks; ++c) {
                    const uint32_t bytes = (uint32_t)c_chunks[c].n * 128u;
                    for (int part = 0; part < parts; ++part) {
                        mbar_wait(bar(kBarEmpty + slot), empty_parity);
                        mbar_expect_tx(bar(kBarFull + slot), bytes);
                        bulk_g2s(s_base + kOffB + slot * kBSlotBytes,
                                 reinterpret_cast<const char *>(a.wimg) + ((size_t)c * 3 + (a.fp16 ? 2 : part)) * kBSlotBytes, bytes, bar(kBarFull + slot));
                        if (++slot == kBSlots) { slot = 0; empty_parity ^= 1u; }
                    }
                }
        }
    } else {
        // =============================================================== MMA issuer: warp-uniform control flow, the elected lane
        // issues (elect_one_sync in tc_common.cuh: keeps the descriptors in uniform registers)
        {
            const uint32_t lead = elect_one_sync();
            constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            auto make_desc = [](uint32_t lo) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi)); return d; };
            const uint32_t a_hi_desc_lo = ((s_base + kOffAHi) >> 4) | (1u << 16), a_lo_desc_lo = ((s_base + kOffALo) >> 4) | (1u << 16);
            const uint32_t b_desc_lo = ((s_base + kOffB) >> 4) | (1u << 16);
            uint32_t slot = 0, full_parity = 0, ready_count = 0;
            auto mma_kblock = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool overwrite_first) {
                umma_bf16_if(lead, d, make_desc(a_lo), make_desc(b_lo), idesc, overwrite_first ? 0u : 1u);
                umma_bf16_if(lead, d, make_desc(a_lo + 2), make_desc(b_lo + 2), idesc, 1u);
                umma_bf16_if(lead, d, make_desc(a_lo + 4), make_desc(b_lo + 4), idesc, 1u);
                umma_bf16_if(lead, d, make_desc(a_lo + 6), make_desc(b_lo + 6), idesc, 1u);
            };
            for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
                int epoch = -1;
                for (int c = 0; c < kNumChunks; ++c) {
                    const NodeChunk ch = c_chunks[c];
                    if ((int)ch.epoch != epoch) {
                        if (epoch >= 0) umma_commit_if(lead, bar(kBarEpochDone));
                        epoch = ch.epoch;
                        mbar_wait(bar(kBarAReady), ready_count & 1u);
                        ++ready_count;
                        tc_fence_after();
                    }
                    const uint32_t idesc = a.fp16 ? idesc_f16(ch.n) : idesc_bf16(ch.n);
                    const uint32_t d = tmem_base + ch.dcol;
                    const uint32_t a_hi = a_hi_desc_lo + ch.a_kb * (kABlockBytes >> 4), a_lo = a_lo_desc_lo + ch.a_kb * (kABlockBytes >> 4);
                    for (int part = 0; part < parts; ++part) {
                        mbar_wait(bar(kBarFull + slot), full_parity);
                        tc_fence_after();
                        const uint32_t b_lo = b_desc_lo + slot * (kBSlotBytes >> 4);
                        if (part == 0) {
                            mma_kblock(d, a_hi, b_lo, idesc, ch.overwrite != 0);
                            if (a.three_pass) mma_kblock(d, a_lo, b_lo, idesc, false);
                        } else {
                            mma_kblock(d, a_hi, b_lo, idesc, false);
                        }
                        umma_commit_if(lead, bar(kBarEmpty + slot));
                        if (++slot == kBSlots) { slot = 0; full_parity ^= 1u; }
                    }
                }
                umma_commit_if(lead, bar(kBarEpochDone));
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

// host-side description of the 34 weight chunks (shared by the packer and the kernel's schedule)
void build_tables(const CnDsrnnWeights &w, PackTable *pack, NodeChunk *chunks)
{
    int c = 0;
    auto add = [&](int epoch, int a_kb, int n, int dcol, bool overwrite, PackSeg s0, PackSeg s1, int nseg) {
        chunks[c] = NodeChunk{(unsigned char)epoch, (unsigned char)a_kb, (unsigned char)(overwrite ? 1 : 0), 0, (unsigned short)n, (unsigned short)dcol};
        pack->c[c].seg[0] = s0; pack->c[c].seg[1] = s1; pack->c[c].nseg = nseg;
        ++c;
    };
    const PackSeg none{nullptr, 0, 0, 0, 0, 0};
    // epochs 0, 1: edge_attention_embed [64, 512]
    for (int kb = 0; kb < 8; ++kb) add(kb / 4, kb % 4, 64, 0, kb == 0, PackSeg{w.n_att_w, 512, 0, kb * 64, 0, 64}, none, 1);
    // epoch 2: node GRU, torch row order r | z | n.  x k-blocks: [n_i | r] -> columns 0..255, [z] -> 256..383;
    //          h k-blocks: [r | z] -> 128..383, [n_h] -> 384..511
    for (int kb = 0; kb < 2; ++kb) {
        add(2, kb, 256, 0, kb == 0, PackSeg{w.n_w_ih, 128, 256, kb * 64, 0, 128}, PackSeg{w.n_w_ih, 128, 0, kb * 64, 128, 128}, 2);
        add(2, kb, 128, 256, kb == 0, PackSeg{w.n_w_ih, 128, 128, kb * 64, 0, 128}, none, 1);
    }
    for (int kb = 0; kb < 2; ++kb) {
        add(2, 2 + kb, 256, 128, false, PackSeg{w.n_w_hh, 128, 0, kb * 64, 0, 256}, none, 1);
        add(2, 2 + kb, 128, 384, kb == 0, PackSeg{w.n_w_hh, 128, 256, kb * 64, 0, 128}, none, 1);
    }
    // epoch 3: output_linear [256, 128]
    for (int kb = 0; kb < 2; ++kb) add(3, kb, 256, 0, kb == 0, PackSeg{w.n_out_w, 128, 0, kb * 64, 0, 256}, none, 1);
    // epoch 4: actor.0 -> columns 0..255, critic.0 -> 256..511
    for (int nb = 0; nb < 2; ++nb)
        for (int kb = 0; kb < 4; ++kb) add(4, kb, 256, nb * 256, kb == 0, PackSeg{nb ? w.critic0_w : w.actor0_w, 256, 0, kb * 64, 0, 256}, none, 1);
    // epochs 5, 6: actor.2, critic.2
    for (int kb = 0; kb < 4; ++kb) add(5, kb, 256, 0, kb == 0, PackSeg{w.actor2_w, 256, 0, kb * 64, 0, 256}, none, 1);
    for (int kb = 0; kb < 4; ++kb) add(6, kb, 256, 0, kb == 0, PackSeg{w.critic2_w, 256, 0, kb * 64, 0, 256}, none, 1);
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host side
void dsrnn_node_tc_destroy(void *state)
{
    NodeTcState *st = static_cast<NodeTcState *>(state);
    if (!st) return;
    cudaFree(st->wimg);
    cudaFree(st->consts);
    delete st;
}

const char *dsrnn_node_tc_repack(void *state, const CnDsrnnWeights *w, cudaStream_t stream);

const char *dsrnn_node_tc_create(const CnDsrnnWeights *w, cudaStream_t stream, void **state)
{
    *state = nullptr;
    NodeTcState *st = new (std::nothrow) NodeTcState();
    if (!st) return "out of host memory";
    st->wimg = nullptr; st->consts = nullptr;
    const size_t img_bytes = (size_t)kNumChunks * 3 * kBSlotBytes;
    if (cudaMalloc(&st->wimg, img_bytes) != cudaSuccess || cudaMalloc(&st->consts, kCTotal * sizeof(float)) != cudaSuccess) {
        dsrnn_node_tc_destroy(st);
        return "cudaMalloc of the packed node / head weights failed";
    }
    PackTable table;
    NodeChunk chunks[kNumChunks];
    build_tables(*w, &table, chunks);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&st->num_sms, cudaDevAttrMultiProcessorCount, dev);
    // the schedule of the 34 chunks does not depend on the weights: uploaded once (`chunks` lives on this stack frame)
    bool ok = cudaMemcpyToSymbolAsync(c_chunks, chunks, sizeof(chunks), 0, cudaMemcpyHostToDevice, stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(stream) == cudaSuccess;
    if (!ok || cudaFuncSetAttribute(node_heads_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024) != cudaSuccess) {
        dsrnn_node_tc_destroy(st);
        return "setting up the node / heads kernel failed";
    }
    if (const char *msg = dsrnn_node_tc_repack(st, w, stream)) { dsrnn_node_tc_destroy(st); return msg; }
    *state = st;
    return nullptr;
}

// packs the current weights into the EXISTING images (stable addresses, no allocation, no synchronisation): the pack
// table (pointers into the caller's parameters) travels as a kernel parameter
const char *dsrnn_node_tc_repack(void *state, const CnDsrnnWeights *w, cudaStream_t stream)
{
    NodeTcState *st = static_cast<NodeTcState *>(state);
    if (!st) return "tensor-core node stage was not initialised";
    PackTable table;
    NodeChunk chunks[kNumChunks];
    build_tables(*w, &table, chunks);
    pack_node_weights_kernel<<<512, 256, 0, stream>>>(table, st->wimg);
    pack_node_consts_kernel<<<8, 256, 0, stream>>>(*w, st->consts);
    return cudaGetLastError() == cudaSuccess ? nullptr : "packing the node / head weights failed";
}

const char *dsrnn_node_tc_forward(void *state, int n_envs, const CnDsrnnIO *io, const float *cat, float *feat, int precision,
                                  cudaStream_t stream, int *launches)
{
    NodeTcState *st = static_cast<NodeTcState *>(state);
    if (!st) return "tensor-core node stage was not initialised";
    NodeTcArgs a;
    a.cat = cat; a.robot_node = io->robot_node; a.h_node_in = io->h_node_in; a.masks = io->masks;
    a.h_node_out = io->h_node_out; a.value = io->value; a.action_mean = io->action_mean; a.feat = feat;
    a.wimg = st->wimg; a.consts = st->consts;
    a.N = n_envs; a.tiles = (n_envs + kRows - 1) / kRows;
    a.three_pass = precision == CN_PREC_BF16X3 ? 1 : 0;
    a.fp16 = precision == CN_PREC_FP16 ? 1 : 0;
    const int grid = a.tiles < st->num_sms ? a.tiles : st->num_sms;
    const cudaError_t lerr = cn_launch(node_heads_tc_kernel, dim3(grid), dim3(kThreads), kSmemBytes + 1024, stream, CN_PDL_NODE, a);
    if (lerr != cudaSuccess) return cudaGetErrorString(lerr);
    ++*launches;
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}
