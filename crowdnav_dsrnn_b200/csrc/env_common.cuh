// env_common.cuh -- device-side state layout, RNG and small helpers shared by the
// crowd-step (K1) and reset (K2) kernels.  sm_100a only.
//
// HBM layout (SoA, one blob owned by the caller, carved by cn_env_create):
//   hum_pv  float4[N*H]  px,py,vx,vy          (read+written every step)
//   hum_gr  float4[N*H]  gx,gy,radius,v_pref  (read every step, written on goal change / reset)
//   hum_bel float4[N*H]  belief px,py,vx,vy   (last_human_states, crowd_sim.py:199,429-455)
//   hum_br  float [N*H]  belief radius
//   hum_th  float [N*H]  spawn heading (only read when a human FOV < 2*pi with a unicycle robot)
//   rob_pv  float4[N]    px,py,vx,vy
//   rob_gr  float4[N]    gx,gy,radius,v_pref
//   rob_x   float4[N]    theta, desiredVelocity[0], potential, episode_return
//   rob_acc float2[N]    last_acceleration
//   ctr     int4  [N]    step_count, scenario_counter, case_counter, current_scenario
// Spare episodes (crowd_reset.cu): the initial state of every env's NEXT episode is generated ahead of time, off the
// critical path, and swapped in by the step kernel when the episode ends:
//   sp_pv / sp_gr / sp_th  [N*H], sp_rob_pv / sp_rob_gr [N], sp_theta [N]   the spawned humans and robot
//   sp_meta int4[N]        valid, case_counter and scenario_counter the spare was generated for, scenario
//   need_spare / need_sync uint8[N], sync_count int[4], refill_list int[N]   work flags / list for the refill and the
//                          synchronous fall-back
//   grp / sp_grp float4[N*CN_MAX_GROUPS]  group environment: the circle groups of the running / spare episode
// Humans of one env are contiguous, envs are contiguous: a CTA that owns E
// consecutive envs reads E*H consecutive float4 (fully coalesced 16 B accesses).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/crowdnav_b200.h"

#define CN_PI 3.141592653589793

struct EnvArrays {
    float4 *hum_pv, *hum_gr, *hum_bel;
    float *hum_br, *hum_th;
    float4 *rob_pv, *rob_gr, *rob_x;
    float2 *rob_acc;
    int4 *ctr;
    float4 *sp_pv, *sp_gr;
    float *sp_th;
    float4 *sp_rob_pv, *sp_rob_gr;
    float *sp_theta;
    int4 *sp_meta;
    uint8_t *need_spare, *need_sync;
    int *sync_count;          // [0] envs flagged in need_sync, [1] CTAs done (sync), [2] entries of refill_list, [3] CTAs done (refill)
    int *refill_list;         // envs whose spare was consumed by the last step (compact: the refill launches a small grid)
    float2 *hum_nv;           // [N*H] new velocities handed from the ORCA kernel to the env-tail kernel (split launch)
    float4 *grp, *sp_grp;     // group environment: [N * CN_MAX_GROUPS] radius, centre x, centre y, valid of the episode's circle groups
};

struct EnvParams {
    CnConfig cfg;
    EnvArrays a;
    int n_envs;
};

static inline size_t cn_align256(size_t x) { return (x + 255) & ~(size_t)255; }

// carve the caller's blob; returns total bytes (base may be NULL to only size it)
static inline size_t cn_carve(EnvArrays *a, void *base, int n, int H)
{
    size_t off = 0;
    char *b = (char *)base;
    const size_t nh = (size_t)n * H;
#define CN_TAKE(field, type, count)                                    \
    do {                                                               \
        if (a) a->field = (type *)(b ? b + off : nullptr);             \
        off = cn_align256(off + sizeof(type) * (count));               \
    } while (0)
    CN_TAKE(hum_pv, float4, nh);
    CN_TAKE(hum_gr, float4, nh);
    CN_TAKE(hum_bel, float4, nh);
    CN_TAKE(hum_br, float, nh);
    CN_TAKE(hum_th, float, nh);
    CN_TAKE(rob_pv, float4, n);
    CN_TAKE(rob_gr, float4, n);
    CN_TAKE(rob_x, float4, n);
    CN_TAKE(rob_acc, float2, n);
    CN_TAKE(ctr, int4, n);
    CN_TAKE(sp_pv, float4, nh);
    CN_TAKE(sp_gr, float4, nh);
    CN_TAKE(sp_th, float, nh);
    CN_TAKE(sp_rob_pv, float4, n);
    CN_TAKE(sp_rob_gr, float4, n);
    CN_TAKE(sp_theta, float, n);
    CN_TAKE(sp_meta, int4, n);
    CN_TAKE(need_spare, uint8_t, n);
    CN_TAKE(need_sync, uint8_t, n);
    CN_TAKE(sync_count, int, 4);
    CN_TAKE(refill_list, int, n);
    CN_TAKE(hum_nv, float2, nh);
    CN_TAKE(grp, float4, (size_t)n * CN_MAX_GROUPS);
    CN_TAKE(sp_grp, float4, (size_t)n * CN_MAX_GROUPS);
#undef CN_TAKE
    return off;
}

// ---------------------------------------------------------------- Philox4x32-10 (counter-based RNG contract)
enum { RNG_RESET = 0, RNG_ATTR = 1, RNG_SPAWN = 2, RNG_GOAL_RANDOM = 3, RNG_GOAL_END = 4, RNG_UNOBS = 5, RNG_POLICY = 6, RNG_GROUP = 7 };
#define RNG_DECISION 0xFFFFFFFFu

__device__ __forceinline__ uint4 philox4x32(uint64_t key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double u01(uint32_t x) { return (double)x * (1.0 / 4294967296.0); }

__device__ __forceinline__ uint64_t episode_key(const CnConfig &cfg, int case_counter, int env_local)
{
    return cfg.seed_offset + (uint64_t)(uint32_t)case_counter + cfg.base_seed +
           (uint64_t)(uint32_t)(cfg.env_id_offset + env_local);
}

// key of the episode an env is IN: reset() used episode_key(case_counter) and then advanced case_counter by nenv
// (crowd_sim_dict.py:162-164), so the running episode's key is one stride back
__device__ __forceinline__ uint64_t running_episode_key(const CnConfig &cfg, int case_counter, int env_local)
{
    const uint64_t cs = cfg.case_size;
    const uint64_t prev = ((uint64_t)(uint32_t)case_counter + cs - (uint64_t)(uint32_t)cfg.nenv % cs) % cs;
    return cfg.seed_offset + prev + cfg.base_seed + (uint64_t)(uint32_t)(cfg.env_id_offset + env_local);
}

// humans.random_policy_changing (crowd_sim.py:463-473): every human of an episode is ORCA or social force with equal
// chance.  The choice is a pure function of (episode key, human), so it is recomputed where needed instead of stored.
__device__ __forceinline__ int human_policy_of(const CnConfig &cfg, int case_counter, int env_local, int i)
{
    if (!cfg.random_policy_changing) return cfg.human_policy;
    return (int)(philox4x32(running_episode_key(cfg, case_counter, env_local), 0u, (uint32_t)i, 0u, RNG_POLICY).x & 1u);
}

__device__ __forceinline__ uint64_t step_key(const CnConfig &cfg, int scenario_counter, int env_local)
{
    return (cfg.base_seed + (uint64_t)(uint32_t)(cfg.env_id_offset + env_local)) ^
           ((uint64_t)(uint32_t)scenario_counter << 32) ^ 0x5EEDC0DE00000000ull;
}

__device__ __forceinline__ double norm2d(double x, double y) { return sqrt(x * x + y * y); }
// exactly (norm2d(x, y) < r), but the fp64 square root (a ~30-instruction Newton sequence) only runs when the squared
// values are within 1e-9 relative of each other; further apart, rounding (1e-16) cannot change the outcome
__device__ __forceinline__ bool norm2d_lt(double x, double y, double r)
{
    const double s = x * x + y * y, r2 = r * r;
    if (!(r > 0.0)) return false;
    if (s < r2 * (1.0 - 1e-9)) return true;
    if (s > r2 * (1.0 + 1e-9)) return false;
    return sqrt(s) < r;
}

__device__ __forceinline__ double shfl_d(unsigned mask, double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(mask, lo, src);
    hi = __shfl_sync(mask, hi, src);
    return __hiloint2double(hi, lo);
}

// detect_visible (crowd_sim/envs/crowd_sim.py:820-847): is (p2) inside the FOV of agent 1.
// heading_src: atan2(vy,vx) for a holonomic-robot config, the agent's theta otherwise.
__device__ __forceinline__ bool detect_visible_d(int kinematics, double p1x, double p1y, double v1x, double v1y,
                                                 double th1, double p2x, double p2y, double fov)
{
    const double theta = (kinematics == CN_HOLONOMIC) ? atan2(v1y, v1x) : th1;
    double fx = cos(theta), fy = sin(theta);
    double dx = p2x - p1x, dy = p2y - p1y;
    const double nf = norm2d(fx, fy);
    fx = fx / nf; fy = fy / nf;
    const double nd = norm2d(dx, dy);
    dx = dx / nd; dy = dy / nd;
    double d = fx * dx + fy * dy;
    if (d != d) return false;
    d = d < -1.0 ? -1.0 : (d > 1.0 ? 1.0 : d);
    return fabs(acos(d)) <= fov / 2.0;
}

// check_collision_group (crowd_sim.py:520-538): does a disc of `radius` at (x, y) touch a circle group (centre distance <=
// group radius + radius + margin; the reference uses 2 * 0.5 for positions / human goals, 4 * 0.5 for the robot goal) or,
// when `movers` is set, a moving (non-obstacle: v_pref != 0) human among the first n_h entries of pv / gr
__device__ __forceinline__ bool collides_with_groups(const float4 *grp, double x, double y, double radius, double margin,
                                                     const float4 *pv, const float4 *gr, int n_h, bool movers)
{
    for (int g = 0; g < CN_MAX_GROUPS; ++g) {
        const float4 q = grp[g];
        if (q.w == 0.0f) break;
        if (norm2d(x - (double)q.y, y - (double)q.z) <= (double)q.x + radius + margin) return true;
    }
    if (movers)
        for (int k = 0; k < n_h; ++k) {
            const float4 g4 = gr[k];
            if (g4.w == 0.0f) continue;                         // isObstacle: static humans have v_pref 0
            const float4 p = pv[k];
            if (norm2d(x - (double)p.x, y - (double)p.y) <= (double)g4.z + radius) return true;
        }
    return false;
}

// observation right after a reset (one warp per env, lane i <-> human i); used by the reset kernel and by the step
// kernel when it swaps a spare episode in
__device__ __forceinline__ void write_reset_obs(const EnvParams &P, const CnObsOut &obs, int e, int lane, int H,
                                                float4 rpv, float4 rgr, float theta, bool reset_flag)
{
    // generate_ob (crowd_sim_dict.py:72-103) from the state in HBM; reset_flag picks the (15,15,0,0,0.3) belief
    const CnConfig &cfg = P.cfg;
    bool vis = false;
    if (lane < H) {
        const size_t hi = (size_t)e * H + lane;
        const float4 hpv = P.a.hum_pv[hi];
        if (cfg.robot_fov >= 2.0 * CN_PI) vis = !((double)hpv.x - (double)rpv.x == 0.0 && (double)hpv.y - (double)rpv.y == 0.0);
        else vis = detect_visible_d(cfg.kinematics, rpv.x, rpv.y, rpv.z, rpv.w, theta, hpv.x, hpv.y, cfg.robot_fov);
        float4 bel;
        if (vis) { bel = hpv; P.a.hum_br[hi] = P.a.hum_gr[hi].z; }
        else if (reset_flag) { bel = make_float4(15.0f, 15.0f, 0.0f, 0.0f); P.a.hum_br[hi] = 0.3f; }
        else {
            bel = P.a.hum_bel[hi];
            bel.x = (float)((double)bel.x + (double)bel.z * cfg.time_step);
            bel.y = (float)((double)bel.y + (double)bel.w * cfg.time_step);
        }
        P.a.hum_bel[hi] = bel;
        if (obs.spatial_edges)
            reinterpret_cast<float2 *>(obs.spatial_edges)[hi] =
                make_float2((float)((double)bel.x - (double)rpv.x), (float)((double)bel.y - (double)rpv.y));
    }
    const unsigned vis_bits = __ballot_sync(0xffffffffu, vis);
    if (lane < 7 && obs.robot_node) {
        const float v = lane == 0 ? rpv.x : lane == 1 ? rpv.y : lane == 2 ? rgr.z : lane == 3 ? rgr.x
                      : lane == 4 ? rgr.y : lane == 5 ? rgr.w : theta;
        obs.robot_node[(size_t)e * 7 + lane] = v;
    } else if (lane >= 7 && lane < 9 && obs.temporal_edges) {
        obs.temporal_edges[(size_t)e * 2 + (lane - 7)] = lane == 7 ? rpv.z : rpv.w;
    }
    if (lane == 0 && obs.visible_mask) obs.visible_mask[e] = vis_bits;
}

// create_agent_attributes (crowd_sim/envs/crowd_sim.py:296-357) from 6 uniforms
struct SpawnCand { double px, py, gx, gy, heading, v_pref; };

__device__ __forceinline__ SpawnCand agent_attributes(const CnConfig &cfg, int scenario, double h_radius, double h_vpref,
                                                      double robot_radius, const double u[6])
{
    SpawnCand c;
    double v_pref = (h_vpref == 0.0) ? 1.0 : h_vpref;
    const double nx = (u[0] - 0.5) * v_pref, ny = (u[1] - 0.5) * v_pref;
    const double R = cfg.circle_radius, sw = cfg.square_width;
    c.heading = 0.0; c.px = c.py = c.gx = c.gy = 0.0;
#define CN_WORLD(uu) (((uu) - 0.5) * sw / 2.0)
    switch (scenario) {
    case CN_SCN_CIRCLE_CROSSING: {
        const double angle = u[2] * CN_PI * 2.0;
        c.px = R * cos(angle) + nx; c.py = R * sin(angle) + ny; c.gx = -c.px; c.gy = -c.py;
    } break;
    case CN_SCN_SQUARE_CROSSING:
        c.px = CN_WORLD(u[2]) * 0.4 + nx; c.py = CN_WORLD(u[3]) * 0.4 + ny;
        c.gx = CN_WORLD(u[4]) * 0.4 + nx; c.gy = CN_WORLD(u[5]) * 0.4 + ny;
        break;
    case CN_SCN_PARALLEL_TRAFFIC: {
        const double sign = (u[2] >= 0.5) ? 1.0 : -1.0;
        c.px = CN_WORLD(u[3]) * 0.4 + nx; c.py = sign * (u[4] * 3.0 + 1.0 + ny); c.gx = c.px; c.gy = -c.py;
    } break;
    case CN_SCN_PERPENDICULAR_TRAFFIC: {
        const double sign = (u[2] >= 0.5) ? 1.0 : -1.0;
        c.px = sign * (u[3] * 3.0 + 1.0 + nx); c.gx = -c.px; c.py = CN_WORLD(u[4]) * 0.4 + ny; c.gy = c.py;
    } break;
    case CN_SCN_SIDE_PREF_PASSING:
    case CN_SCN_SIDE_PREF_OVERTAKING: {
        const double min_x = -(robot_radius + h_radius), max_x = -min_x;
        const double hx = (max_x - min_x) * u[2] + min_x;
        c.px = hx; c.gx = hx;
        if (scenario == CN_SCN_SIDE_PREF_PASSING) { c.py = R; c.gy = -R; c.heading = -CN_PI / 2.0; }
        else { c.py = -R + 2.0; c.gy = R + 2.0; c.heading = CN_PI / 2.0; v_pref = 0.3; }
    } break;
    case CN_SCN_SIDE_PREF_CROSSING: {
        const double min_x = -(R + robot_radius + h_radius), max_x = -(R - robot_radius - h_radius);
        const double hx = (max_x - min_x) * u[2] + min_x;
        c.px = hx; c.gx = -hx; c.py = 0.0; c.gy = 0.0;
    } break;
    default: break;
    }
#undef CN_WORLD
    c.v_pref = v_pref;
    return c;
}
