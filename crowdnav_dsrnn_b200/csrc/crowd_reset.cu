// crowd_reset.cu -- K2: device-side episode reset (sm_100a), plus the observation-only kernel and the
// canonical <-> SoA state converters used by the parity tests.  Compile with -fmad=false.
//
// CrowdSimDict.reset (crowd_sim/envs/crowd_sim_dict.py:105-203): scenario choice, robot spawn
// (crowd_sim.py:626-660), per-human attributes (agent.py:44-50) and spawn/goal by scenario
// (crowd_sim.py:296-393) with the min-distance rejection rule, belief initialisation
// (crowd_sim.py:443-445), potential, counters.  One CTA per resetting env; thread t evaluates try t of the
// (bounded) rejection loops, the first accepted thread wins -- identical to the sequential loop
// because every candidate is a pure function of (episode key, human, try) under the counter-based
// Philox4x32-10 contract (oracle/crowd_oracle.c restates the same contract sequentially).
#include <cstdio>
#include <cstdlib>
#include "env_common.cuh"

#define RESET_THREADS 256

// first thread (in try order) whose candidate is acceptable, over the whole CTA; -1 if none.  s_vote: 4 ints.
__device__ __forceinline__ int cta_first_ok(bool ok, int *s_vote)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    __syncthreads();                       // previous round's readers are done with s_vote
    if (lane == 0) s_vote[warp] = b ? warp * 32 + (__ffs(b) - 1) : 0x7fffffff;
    __syncthreads();
    int first = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) first = min(first, s_vote[w]);
    return first == 0x7fffffff ? -1 : first;
}

// generate_robot_humans of the GROUP environment (sim.group_human, crowd_sim.py:559-622): circles of 4-9 static humans
// (generate_circle_group_obstacle, :476-518) around random centres in [-3, 3]^2 until at most 4 humans are left, those walk
// (generate_circle_crossing_human with the group collision rule, :371-372); the robot starts on the circle of radius 5.5 at a
// random angle and its goal lies opposite, both stepped by 0.2 rad past the groups.  Same CTA-parallel bounded tries and Philox
// contract as the plain spawn (oracle/crowd_oracle.c restates it sequentially).
__device__ __noinline__ void reset_group_env(const EnvParams &P, int e, uint64_t key, int scenario, float4 *dst_pv, float4 *dst_gr,
                                             float *dst_th, float4 *dst_grp, int *s_vote, double *s_pick, float4 &rpv_out, float4 &rgr_out)
{
    __shared__ float4 s_grp[CN_MAX_GROUPS];
    __shared__ float4 s_pv[CN_MAX_HUMANS], s_gr[CN_MAX_HUMANS];
    const CnConfig &cfg = P.cfg;
    const int H = cfg.human_num, tid = threadIdx.x;
    __syncthreads();
    if (tid < CN_MAX_GROUPS) s_grp[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    int left = H, idx = 0, ng = 0;
    while (left > 0) {
        if (left <= 4) {
            for (; idx < H; ++idx) {
                double v_pref = cfg.human_v_pref, radius = cfg.human_radius;
                if (cfg.randomize_attributes) {
                    const uint4 x = philox4x32(key, 0, (uint32_t)idx, 0, RNG_ATTR);
                    v_pref = 0.5 + (1.5 - 0.5) * u01(x.x);
                    radius = 0.3 + (0.5 - 0.3) * u01(x.y);
                }
                const float radius_f = (float)radius;
                SpawnCand c;
                c.px = c.py = c.gx = c.gy = c.heading = c.v_pref = 0.0;
                for (int t0 = 0; t0 < cfg.max_spawn_tries; t0 += (int)blockDim.x) {
                    const int t = t0 + tid;
                    const uint4 xa = philox4x32(key, (uint32_t)t, (uint32_t)idx, 0, RNG_SPAWN);
                    const uint4 xb = philox4x32(key, (uint32_t)t, (uint32_t)idx, 1, RNG_SPAWN);
                    const double u6[6] = {u01(xa.x), u01(xa.y), u01(xa.z), u01(xa.w), u01(xb.x), u01(xb.y)};
                    c = agent_attributes(cfg, scenario, (double)radius_f, v_pref, (double)(float)cfg.robot_radius, u6);
                    const bool ok = t < cfg.max_spawn_tries &&
                                    !collides_with_groups(s_grp, c.px, c.py, (double)radius_f, 2 * 0.5, s_pv, s_gr, idx, true);
                    int src = cta_first_ok(ok, s_vote);
                    if (src < 0 && t0 + (int)blockDim.x >= cfg.max_spawn_tries) src = (cfg.max_spawn_tries - 1) - t0;     // keep the last try
                    if (src >= 0) {
                        if (tid == src) { s_pick[0] = c.px; s_pick[1] = c.py; s_pick[2] = c.gx; s_pick[3] = c.gy; s_pick[4] = c.heading; s_pick[5] = c.v_pref; }
                        __syncthreads();
                        c.px = s_pick[0]; c.py = s_pick[1]; c.gx = s_pick[2]; c.gy = s_pick[3]; c.heading = s_pick[4]; c.v_pref = s_pick[5];
                        break;
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    const size_t hi = (size_t)e * H + idx;
                    s_pv[idx] = make_float4((float)c.px, (float)c.py, 0.0f, 0.0f);
                    s_gr[idx] = make_float4((float)c.gx, (float)c.gy, radius_f, (float)c.v_pref);
                    dst_pv[hi] = s_pv[idx]; dst_gr[hi] = s_gr[idx]; dst_th[hi] = (float)c.heading;
                }
                __syncthreads();
            }
            left = 0;
        } else {
            const int max_rand = left < 10 ? left : 10;
            int circum = 4 + (int)(u01(philox4x32(key, (uint32_t)ng, 0, 0, RNG_GROUP).x) * (double)(max_rand - 4));     // randint(4, max_rand)
            if (circum > max_rand - 1) circum = max_rand - 1;
            const double g_radius = cfg.human_radius * 2.0 * circum / (2.0 * CN_PI);
            double cx = 0.0, cy = 0.0;
            for (int t0 = 0; t0 < cfg.max_spawn_tries; t0 += (int)blockDim.x) {
                const int t = t0 + tid;
                const uint4 x = philox4x32(key, (uint32_t)t, (uint32_t)ng, 1, RNG_GROUP);
                cx = -3.0 + 6.0 * u01(x.x); cy = -3.0 + 6.0 * u01(x.y);
                bool ok = t < cfg.max_spawn_tries;
                for (int g = 0; g < ng; ++g) {
                    const float4 q = s_grp[g];
                    if (norm2d(cx - (double)q.y, cy - (double)q.z) < g_radius + (double)q.x + 2.0 * cfg.human_radius) ok = false;
                }
                int src = cta_first_ok(ok, s_vote);
                if (src < 0 && t0 + (int)blockDim.x >= cfg.max_spawn_tries) src = (cfg.max_spawn_tries - 1) - t0;
                if (src >= 0) {
                    if (tid == src) { s_pick[0] = cx; s_pick[1] = cy; }
                    __syncthreads();
                    cx = s_pick[0]; cy = s_pick[1];
                    break;
                }
            }
            __syncthreads();
            if (tid == 0) s_grp[ng] = make_float4((float)g_radius, (float)cx, (float)cy, 1.0f);
            __syncthreads();
            if (tid < circum) {
                const float4 q = s_grp[ng];
                const double angle = (2.0 * CN_PI / circum) * tid;
                const float px = (float)((double)q.y + (double)q.x * cos(angle)), py = (float)((double)q.z + (double)q.x * sin(angle));
                const size_t hi = (size_t)e * H + idx + tid;
                s_pv[idx + tid] = make_float4(px, py, 0.0f, 0.0f);
                s_gr[idx + tid] = make_float4(px, py, (float)cfg.human_radius, 0.0f);      // static: goal = position, v_pref = 0
                dst_pv[hi] = s_pv[idx + tid]; dst_gr[hi] = s_gr[idx + tid]; dst_th[hi] = 0.0f;
            }
            __syncthreads();
            idx += circum; left -= circum; ++ng;
        }
    }
    if (tid < CN_MAX_GROUPS) dst_grp[tid] = s_grp[tid];
    // robot start / goal: thread t evaluates the t-th 0.2 rad increment (accumulated like the reference's `+= 0.2`)
    const double rand_angle = u01(philox4x32(key, 0, 0, 2, RNG_GROUP).x) * CN_PI * 2.0;
    double inc = 0.0;
    for (int k = 0; k < tid && k < 63; ++k) inc = inc + 0.2;
    double px = cos(rand_angle + inc) * 5.5, py = sin(rand_angle + inc) * 5.5;
    {
        const bool ok = tid < 64 && !collides_with_groups(s_grp, px, py, cfg.robot_radius, 2 * 0.5, s_pv, s_gr, H, true);
        int src = cta_first_ok(ok, s_vote);
        const bool none = src < 0;
        if (none) src = 63;
        if (tid == src) { s_pick[0] = px; s_pick[1] = py; s_pick[2] = none ? inc + 0.2 : inc; }
        __syncthreads();
        px = s_pick[0]; py = s_pick[1]; inc = s_pick[2];
        __syncthreads();
    }
    inc = inc + CN_PI;
    for (int k = 0; k < tid && k < 63; ++k) inc = inc + 0.2;
    double gx = cos(rand_angle + inc) * 5.5, gy = sin(rand_angle + inc) * 5.5;
    {
        const bool ok = tid < 64 && !collides_with_groups(s_grp, gx, gy, cfg.robot_radius, 4 * 0.5, s_pv, s_gr, H, false);
        int src = cta_first_ok(ok, s_vote);
        if (src < 0) src = 63;
        if (tid == src) { s_pick[0] = gx; s_pick[1] = gy; }
        __syncthreads();
        gx = s_pick[0]; gy = s_pick[1];
        __syncthreads();
    }
    rpv_out = make_float4((float)px, (float)py, 0.0f, 0.0f);
    rgr_out = make_float4((float)gx, (float)gy, (float)cfg.robot_radius, (float)cfg.robot_v_pref);
}

// grid: one CTA of RESET_THREADS threads per env (envs whose mask byte is 0 exit at once); thread t evaluates try
// t, t+RESET_THREADS, ... of each bounded rejection loop, so a typical spawn needs a single round per human.
// (-DRESET_PROFILE prints per-env cycle counts of the phases; tools/reset_stats.py summarises them.)
//
// mode CN_RESET_LIVE   reset the envs whose mask byte is set (all when mask == NULL): cn_env_reset
// mode CN_RESET_SPARE  generate the NEXT episode of every env flagged in need_spare into the spare arrays, from the
//                      counters the next reset will see; runs on a side stream, off the critical path
// mode CN_RESET_SYNC   fall-back of the step kernel: envs flagged in need_sync (their spare was missing or stale, e.g.
//                      after cn_env_set_state changed the counters); exits at once when sync_count is 0
// mode CN_RESET_SPARE_LIST  like CN_RESET_SPARE for the envs in refill_list (appended by the step kernel / the fall-back):
//                      a small grid instead of one (mostly idle) CTA per env
enum { CN_RESET_LIVE = 0, CN_RESET_SPARE = 1, CN_RESET_SYNC = 2, CN_RESET_SPARE_LIST = 3 };

template <bool kGroup>      // kGroup: the group environment (its own instantiation: the plain reset keeps its registers)
__global__ void __launch_bounds__(RESET_THREADS)
crowd_reset_kernel(const __grid_constant__ EnvParams P, const __grid_constant__ CnObsOut obs,
                   const uint8_t *mask, int mode)
{
    __shared__ double s_hx[CN_MAX_HUMANS], s_hy[CN_MAX_HUMANS], s_hr[CN_MAX_HUMANS];   // px, py, radius of the humans spawned so far
    // fp32 pre-filter of the min-distance rule against earlier human k: squared distance below s_lo[k] / above s_hi[k]
    // decides the fp64 comparison for certain (band of 1e-4 relative, fp32 error of the squared distance < 1e-5 there)
    __shared__ float s_fx[CN_MAX_HUMANS], s_fy[CN_MAX_HUMANS], s_lo[CN_MAX_HUMANS], s_hi[CN_MAX_HUMANS];
    __shared__ int s_vote[RESET_THREADS / 32];
    __shared__ double s_pick[6];
    const CnConfig &cfg = P.cfg;
    const int H = cfg.human_num;
    const int tid = threadIdx.x, lane = threadIdx.x & 31;
    const bool spare = mode == CN_RESET_SPARE || mode == CN_RESET_SPARE_LIST;
    const int *list = nullptr;
    int limit = P.n_envs;
    if (mode == CN_RESET_SYNC) {
        if (P.a.sync_count[0] == 0) return;
        mask = P.a.need_sync;
    } else if (mode == CN_RESET_SPARE_LIST) {
        list = P.a.refill_list;
        limit = min(P.a.sync_count[2], P.n_envs);
        mask = nullptr;
    } else if (spare) {
        mask = P.a.need_spare;
    }
    float4 *dst_pv = spare ? P.a.sp_pv : P.a.hum_pv, *dst_gr = spare ? P.a.sp_gr : P.a.hum_gr;
    float *dst_th = spare ? P.a.sp_th : P.a.hum_th;
    for (int it = blockIdx.x; it < limit; it += gridDim.x) {      // one env per CTA in the full-grid modes
    const int e = list ? list[it] : it;
    if (mask && !mask[e]) continue;
#ifdef RESET_PROFILE
    const long long pt0 = clock64();
    long long pt_mid = 0;
    long long pq[3] = {0, 0, 0};
    int prounds = 0;
#endif
    int4 ctr = P.a.ctr[e];
    const uint64_t key = episode_key(cfg, ctr.z, e);
    const uint4 g0 = philox4x32(key, 0, 0, 0, RNG_RESET);
    int scn_idx;
    if (cfg.social_metrics) scn_idx = (int)((uint32_t)ctr.y % 4u);
    else { scn_idx = (int)(u01(g0.x) * cfg.n_scenarios); if (scn_idx >= cfg.n_scenarios) scn_idx = cfg.n_scenarios - 1; }
    const int scenario = cfg.scenarios[scn_idx];
    const double R = cfg.circle_radius;

    float4 rpv, rgr;
    double rth;
    if (kGroup) {
        // ---- the group environment spawns its humans first and the robot around them (crowd_sim.py:559-622)
        reset_group_env(P, e, key, scenario, dst_pv, dst_gr, dst_th, (spare ? P.a.sp_grp : P.a.grp) + (size_t)e * CN_MAX_GROUPS,
                        s_vote, s_pick, rpv, rgr);
        rth = CN_PI / 2.0;
    } else {
    // ---- robot (crowd_sim.py:626-660)
    double rpx, rpy, rgx = 0.0, rgy = 0.0;
    if (cfg.kinematics == CN_UNICYCLE || !(cfg.social_metrics || cfg.side_preference)) {
        const bool uni = cfg.kinematics == CN_UNICYCLE;
        const double angle = u01(g0.y) * CN_PI * 2.0;
        double cpx = R * cos(angle), cpy = R * sin(angle), cgx = 0.0, cgy = 0.0;
        for (int t0 = 0; t0 < cfg.max_robot_tries; t0 += (int)blockDim.x) {
            const int t = t0 + tid;
            const uint4 x = philox4x32(key, (uint32_t)t, 1, 0, RNG_RESET);
            if (uni) { cgx = -R + 2.0 * R * u01(x.x); cgy = -R + 2.0 * R * u01(x.y); }
            else {
                cpx = -R + 2.0 * R * u01(x.x); cpy = -R + 2.0 * R * u01(x.y);
                cgx = -R + 2.0 * R * u01(x.z); cgy = -R + 2.0 * R * u01(x.w);
            }
            const bool ok = t < cfg.max_robot_tries && !norm2d_lt(cpx - cgx, cpy - cgy, 6.0);
            int src = cta_first_ok(ok, s_vote);
            const bool last_round = t0 + (int)blockDim.x >= cfg.max_robot_tries;
            if (src < 0 && last_round) src = (cfg.max_robot_tries - 1) - t0;        // keep the last try
            if (src >= 0) {
                if (tid == src) { s_pick[0] = cpx; s_pick[1] = cpy; s_pick[2] = cgx; s_pick[3] = cgy; }
                __syncthreads();
                cpx = s_pick[0]; cpy = s_pick[1]; cgx = s_pick[2]; cgy = s_pick[3];
                break;
            }
        }
        rpx = cpx; rpy = cpy; rgx = cgx; rgy = cgy;
        rth = uni ? u01(g0.z) * 2.0 * CN_PI : CN_PI / 2.0;
    } else {
        rpx = 0.0; rpy = -R; rgx = 0.0; rgy = R; rth = CN_PI / 2.0;
    }
    rpv = make_float4((float)rpx, (float)rpy, 0.0f, 0.0f);
    rgr = make_float4((float)rgx, (float)rgy, (float)cfg.robot_radius, (float)cfg.robot_v_pref);

#ifdef RESET_PROFILE
    const long long pt1 = clock64();
#endif
    // ---- humans, sequentially; tries in parallel (crowd_sim.py:359-393)
    for (int i = 0; i < H; ++i) {
#ifdef RESET_PROFILE
        if (i == H / 2) pt_mid = clock64();
#endif
        double v_pref = cfg.human_v_pref, radius = cfg.human_radius;
        if (cfg.randomize_attributes) {
            const uint4 x = philox4x32(key, 0, (uint32_t)i, 0, RNG_ATTR);
            v_pref = 0.5 + (1.5 - 0.5) * u01(x.x);
            radius = 0.3 + (0.5 - 0.3) * u01(x.y);
        }
        const float radius_f = (float)radius;
        if (tid < i) {
            const double md = (double)radius_f + s_hr[tid] + cfg.discomfort_dist, md2 = md * md;
            s_lo[tid] = (float)(md2 * (1.0 - 1e-4));
            s_hi[tid] = (float)(md2 * (1.0 + 1e-4));
        }
        __syncthreads();
        SpawnCand c;
        c.px = c.py = c.gx = c.gy = c.heading = c.v_pref = 0.0;
        for (int t0 = 0; t0 < cfg.max_spawn_tries; t0 += (int)blockDim.x) {
            const int t = t0 + tid;
#ifdef RESET_PROFILE
            ++prounds;
            const long long q0 = clock64();
#endif
            const uint4 xa = philox4x32(key, (uint32_t)t, (uint32_t)i, 0, RNG_SPAWN);
            const uint4 xb = philox4x32(key, (uint32_t)t, (uint32_t)i, 1, RNG_SPAWN);
            const double u6[6] = {u01(xa.x), u01(xa.y), u01(xa.z), u01(xa.w), u01(xb.x), u01(xb.y)};
            c = agent_attributes(cfg, scenario, (double)radius_f, v_pref, (double)rgr.z, u6);
#ifdef RESET_PROFILE
            const long long q1 = clock64();
#endif
            bool collide;
            {
                const double md = (cfg.kinematics == CN_UNICYCLE) ? R / 2.0 : (double)radius_f + (double)rgr.z + cfg.discomfort_dist;
                collide = norm2d_lt(c.px - (double)rpv.x, c.py - (double)rpv.y, md);
            }
            // branch-free fp32 pass over the earlier humans (positions are stored as floats, so s_fx / s_fy are exact);
            // the exact fp64 rule only runs for a candidate that lands inside some pair's 1e-4 band without a certain hit
            {
                const float fx = (float)c.px, fy = (float)c.py;
                bool band = false;
#pragma unroll 4
                for (int k = 0; k < i; ++k) {
                    const float dx = fx - s_fx[k], dy = fy - s_fy[k];
                    const float sq = dx * dx + dy * dy;
                    collide |= sq < s_lo[k];
                    band |= (sq >= s_lo[k]) & (sq <= s_hi[k]);
                }
                if (band && !collide) {
                    for (int k = 0; k < i; ++k) {
                        const double md = (double)radius_f + s_hr[k] + cfg.discomfort_dist;
                        collide |= norm2d_lt(c.px - s_hx[k], c.py - s_hy[k], md);
                    }
                }
            }
            const bool ok = t < cfg.max_spawn_tries && !collide;
#ifdef RESET_PROFILE
            const long long q2 = clock64();
#endif
            int src = cta_first_ok(ok, s_vote);
#ifdef RESET_PROFILE
            const long long q3 = clock64();
            pq[0] += q1 - q0; pq[1] += q2 - q1; pq[2] += q3 - q2;
#endif
            const bool last_round = t0 + (int)blockDim.x >= cfg.max_spawn_tries;
            if (src < 0 && last_round) src = (cfg.max_spawn_tries - 1) - t0;        // keep the last try
            if (src >= 0) {
                if (tid == src) { s_pick[0] = c.px; s_pick[1] = c.py; s_pick[2] = c.gx; s_pick[3] = c.gy; s_pick[4] = c.heading; s_pick[5] = c.v_pref; }
                __syncthreads();
                c.px = s_pick[0]; c.py = s_pick[1]; c.gx = s_pick[2]; c.gy = s_pick[3]; c.heading = s_pick[4]; c.v_pref = s_pick[5];
                break;
            }
        }
        const float4 pv = make_float4((float)c.px, (float)c.py, 0.0f, 0.0f);
        const float4 gr = make_float4((float)c.gx, (float)c.gy, radius_f, (float)c.v_pref);
        __syncthreads();                    // everyone has read s_pick / the old s_h before they change
        if (tid == 0) {
            const size_t hi = (size_t)e * H + i;
            s_hx[i] = (double)pv.x; s_hy[i] = (double)pv.y; s_hr[i] = (double)radius_f;
            s_fx[i] = pv.x; s_fy[i] = pv.y;
            dst_pv[hi] = pv;
            dst_gr[hi] = gr;
            dst_th[hi] = (float)c.heading;
        }
        __syncthreads();
    }
    }   // !group_human

#ifdef RESET_PROFILE
    if (tid == 0 && e < 1500) printf("reset e=%d scn=%d robot %lld first-half %lld second-half %lld cycles, %d rounds | draw %lld collide %lld vote %lld\n", e, scenario,
                                     pt1 - pt0, pt_mid - pt1, clock64() - pt_mid, prounds, pq[0], pq[1], pq[2]);
#endif
    // ---- counters, potential, observation (warp 0 only; its lane 0 wrote the humans above)
    if (tid < 32) {
        if (spare) {
            if (lane == 0) {
                P.a.sp_rob_pv[e] = rpv;
                P.a.sp_rob_gr[e] = rgr;
                P.a.sp_theta[e] = (float)rth;
                P.a.sp_meta[e] = make_int4(1, ctr.z, ctr.y, scenario);
                P.a.need_spare[e] = 0;
            }
        } else {
            ctr.x = 0;
            ctr.z = (int)(uint32_t)(((uint64_t)(uint32_t)ctr.z + (uint64_t)cfg.nenv) % cfg.case_size);
            ctr.w = scenario;
            ctr.y += 1;
            __threadfence_block();
            write_reset_obs(P, obs, e, lane, H, rpv, rgr, (float)rth, true);
            if (lane == 0) {
                float4 rx = P.a.rob_x[e];
                rx.x = (float)rth;
                rx.y = 0.0f;
                rx.z = (float)(-fabs(norm2d((double)rpv.x - (double)rgr.x, (double)rpv.y - (double)rgr.y)));
                rx.w = 0.0f;
                P.a.rob_pv[e] = rpv;
                P.a.rob_gr[e] = rgr;
                P.a.rob_x[e] = rx;
                P.a.ctr[e] = ctr;
                P.a.sp_meta[e] = make_int4(0, 0, 0, 0);     // the counters moved on: whatever spare there was is stale
                if (mode == CN_RESET_SYNC) {
                    P.a.need_sync[e] = 0;
                    const int slot = atomicAdd(&P.a.sync_count[2], 1);
                    if (slot < P.n_envs) P.a.refill_list[slot] = e;
                } else {
                    P.a.need_spare[e] = 1;
                }
            }
        }
    }
    __syncthreads();          // shared arrays are reused by the next env of this CTA (fall-back mode)
    }
    if (mode == CN_RESET_SYNC) {          // the last CTA to finish re-arms the counter for the next step
        __threadfence();
        if (tid == 0 && atomicAdd(&P.a.sync_count[1], 1) == (int)gridDim.x - 1) { P.a.sync_count[0] = 0; P.a.sync_count[1] = 0; }
    } else if (mode == CN_RESET_SPARE_LIST) {
        __threadfence();
        if (tid == 0 && atomicAdd(&P.a.sync_count[3], 1) == (int)gridDim.x - 1) { P.a.sync_count[2] = 0; P.a.sync_count[3] = 0; }
    }
}

__global__ void __launch_bounds__(RESET_THREADS)
crowd_observe_kernel(const __grid_constant__ EnvParams P, const __grid_constant__ CnObsOut obs)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * (RESET_THREADS / 32) + warp;
    if (e >= P.n_envs) return;
    write_reset_obs(P, obs, e, lane, P.cfg.human_num, P.a.rob_pv[e], P.a.rob_gr[e], P.a.rob_x[e].x, true);
}

// canonical [N,9]/[N,H,9]/[N,H,5]/[N,4]/[N,4]/[N] view  <->  SoA   (dir 0: view -> SoA, 1: SoA -> view)
__global__ void state_convert_kernel(const __grid_constant__ EnvParams P, const __grid_constant__ CnStateView v, int dir)
{
    const int H = P.cfg.human_num;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nh = (size_t)P.n_envs * H;
    if (idx < nh) {
        if (v.humans) {
            float *h = v.humans + idx * 9;
            if (dir == 0) {
                P.a.hum_pv[idx] = make_float4(h[0], h[1], h[2], h[3]);
                P.a.hum_gr[idx] = make_float4(h[5], h[6], h[4], h[7]);
                P.a.hum_th[idx] = h[8];
            } else {
                const float4 pv = P.a.hum_pv[idx], gr = P.a.hum_gr[idx];
                h[0] = pv.x; h[1] = pv.y; h[2] = pv.z; h[3] = pv.w; h[4] = gr.z; h[5] = gr.x; h[6] = gr.y; h[7] = gr.w;
                h[8] = P.a.hum_th[idx];
            }
        }
        if (v.belief) {
            float *b = v.belief + idx * 5;
            if (dir == 0) { P.a.hum_bel[idx] = make_float4(b[0], b[1], b[2], b[3]); P.a.hum_br[idx] = b[4]; }
            else { const float4 q = P.a.hum_bel[idx]; b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w; b[4] = P.a.hum_br[idx]; }
        }
    }
    if (idx < (size_t)P.n_envs) {
        float4 rx = P.a.rob_x[idx];
        if (v.robot) {
            float *r = v.robot + idx * 9;
            if (dir == 0) {
                P.a.rob_pv[idx] = make_float4(r[0], r[1], r[2], r[3]);
                P.a.rob_gr[idx] = make_float4(r[5], r[6], r[4], r[7]);
                rx.x = r[8];
            } else {
                const float4 pv = P.a.rob_pv[idx], gr = P.a.rob_gr[idx];
                r[0] = pv.x; r[1] = pv.y; r[2] = pv.z; r[3] = pv.w; r[4] = gr.z; r[5] = gr.x; r[6] = gr.y; r[7] = gr.w; r[8] = rx.x;
            }
        }
        if (v.extras) {
            float *x = v.extras + idx * 4;
            if (dir == 0) { rx.y = x[0]; rx.z = x[1]; P.a.rob_acc[idx] = make_float2(x[2], x[3]); }
            else { const float2 a = P.a.rob_acc[idx]; x[0] = rx.y; x[1] = rx.z; x[2] = a.x; x[3] = a.y; }
        }
        if (v.episode_return) {
            if (dir == 0) rx.w = v.episode_return[idx]; else v.episode_return[idx] = rx.w;
        }
        if (dir == 0) P.a.rob_x[idx] = rx;
        if (v.groups) {
            for (int g = 0; g < CN_MAX_GROUPS; ++g) {
                float *q = v.groups + (idx * CN_MAX_GROUPS + g) * 4;
                float4 &d = P.a.grp[idx * CN_MAX_GROUPS + g];
                if (dir == 0) d = make_float4(q[0], q[1], q[2], q[3]);
                else { q[0] = d.x; q[1] = d.y; q[2] = d.z; q[3] = d.w; }
            }
        }
        if (v.counters) {
            int32_t *c = v.counters + idx * 4;
            if (dir == 0) P.a.ctr[idx] = make_int4(c[0], c[1], c[2], c[3]);
            else { const int4 q = P.a.ctr[idx]; c[0] = q.x; c[1] = q.y; c[2] = q.z; c[3] = q.w; }
        }
    }
}

extern "C" int cn_launch_crowd_reset(const EnvParams *P, const CnObsOut *obs, const uint8_t *mask, int mode, cudaStream_t stream)
{
    CnObsOut none = {};
    // the fall-back normally finds nothing to do: a small grid that strides over the envs keeps its launch cheap
    const int small = mode == CN_RESET_SPARE_LIST ? 1184 : 592;
    const int grid = (mode == CN_RESET_SYNC || mode == CN_RESET_SPARE_LIST) ? (P->n_envs < small ? P->n_envs : small) : P->n_envs;
    // The refill of the spare episodes runs beside the forward's attention kernel: smaller CTAs (thread t = try t, t + threads, ...;
    // a spawn rarely needs more than a few tries) keep less of an SM away from it.  Same candidates, same first accepted try.
    int threads = RESET_THREADS;
    if (!P->cfg.group_human && (mode == CN_RESET_SPARE || mode == CN_RESET_SPARE_LIST)) {
        threads = 64;
        if (const char *dbg = getenv("CN_REFILL_THREADS")) { const int v = atoi(dbg); if (v == 32 || v == 64 || v == 128 || v == 256) threads = v; }
    }
    if (P->cfg.group_human) crowd_reset_kernel<true><<<grid, threads, 0, stream>>>(*P, obs ? *obs : none, mask, mode);
    else crowd_reset_kernel<false><<<grid, threads, 0, stream>>>(*P, obs ? *obs : none, mask, mode);
    return (int)cudaGetLastError();
}

extern "C" int cn_launch_crowd_observe(const EnvParams *P, const CnObsOut *obs, cudaStream_t stream)
{
    const int per = RESET_THREADS / 32;
    crowd_observe_kernel<<<(P->n_envs + per - 1) / per, RESET_THREADS, 0, stream>>>(*P, *obs);
    return (int)cudaGetLastError();
}

extern "C" int cn_launch_state_convert(const EnvParams *P, const CnStateView *v, int dir, cudaStream_t stream)
{
    const size_t total = (size_t)P->n_envs * P->cfg.human_num;
    const int threads = 256;
    state_convert_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(*P, *v, dir);
    return (int)cudaGetLastError();
}
