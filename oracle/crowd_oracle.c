/*
 * oracle/crowd_oracle.c -- TEST INFRASTRUCTURE (CPU oracle). Not product code.
 *
 * Sequential plain-C restatement of the reference's crowd step for ONE env at
 * a time (looped / OpenMP'd over envs):
 *
 *   clip_action            crowd_nav/policy/srnn.py:18-48
 *   unicycle accumulation  crowd_sim/envs/crowd_sim_dict.py:211-217
 *   get_human_actions      crowd_sim/envs/crowd_sim.py:1121-1161  (+ ORCA.predict, crowd_nav/policy/orca.py:64-139;
 *                          SOCIAL_FORCE.predict, crowd_nav/policy/social_force.py:11-63; random_unobservability;
 *                          random_policy_changing, crowd_sim.py:463-473)
 *   detect_visible         crowd_sim/envs/crowd_sim.py:820-847
 *   calc_reward            crowd_sim/envs/crowd_sim.py:907-1094
 *   Agent.step             crowd_sim/envs/utils/agent.py:172-212
 *   generate_ob            crowd_sim/envs/crowd_sim_dict.py:72-103, crowd_sim.py:429-455, 851-865
 *   goal updates           crowd_sim/envs/crowd_sim.py:724-811, crowd_sim_dict.py:261-269
 *   reset                  crowd_sim/envs/crowd_sim_dict.py:105-203, crowd_sim.py:296-393, 555-663
 *   group environment      crowd_sim/envs/crowd_sim.py:476-622 (circle groups of static humans, sim.group_human)
 *
 * Pinning: the deterministic part (everything except the RNG-driven reset and
 * goal re-sampling) is checked against the reference's own Python executed in
 * the build container on injected states (oracle/gen_golden.py ->
 * tests/golden/step_*.npz).  ORCA is the restated RVO2 of orca_core.h
 * (PARITY UNPINNED against the real rvo2 wheel, which is not installable).
 * The RNG-driven parts use a counter-based Philox4x32-10 instead of the
 * reference's global MT19937 stream and BOUNDED rejection loops (documented
 * deviations); they are pinned only distributionally.
 *
 * State is float32 (the product's HBM layout); every non-ORCA expression is
 * evaluated in double like the reference's Python floats.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "crowd_oracle.h"
#include "orca_core.h"

#define PI 3.141592653589793

/* ------------------------------------------------------------------ Philox4x32-10 */
void oracle_philox(uint64_t key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void philox_u01(uint64_t key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double u[4])
{
    uint32_t x[4];
    oracle_philox(key, c0, c1, c2, c3, x);
    for (int i = 0; i < 4; ++i) u[i] = (double)x[i] * (1.0 / 4294967296.0);
}

enum { RNG_RESET = 0, RNG_ATTR = 1, RNG_SPAWN = 2, RNG_GOAL_RANDOM = 3, RNG_GOAL_END = 4, RNG_UNOBS = 5, RNG_POLICY = 6, RNG_GROUP = 7 };
#define RNG_DECISION 0xFFFFFFFFu

/* ------------------------------------------------------------------ per-env views */
typedef struct {
    const CnConfig *cfg;
    int H;
    float *rob;   /* 9 */
    float *hum;   /* H*9 */
    float *bel;   /* H*5 */
    float *ext;   /* 4: desired_v, potential, last_ax, last_ay */
    int32_t *ctr; /* 4: step_count, scenario_counter, case_counter, scenario */
    float *ep_ret;
    float *grp;   /* CN_MAX_GROUPS * 4: radius, cx, cy, valid */
    uint64_t env_gid;
} Env;

enum { PX = 0, PY = 1, VX = 2, VY = 3, RAD = 4, GX = 5, GY = 6, VPREF = 7, TH = 8 };

static double norm2(double x, double y) { return sqrt(x * x + y * y); }

/* detect_visible (crowd_sim.py:820-847): is agent 2 inside agent 1's FOV */
static int detect_visible(const CnConfig *cfg, const float *a1, const float *a2, double fov)
{
    double theta;
    if (cfg->kinematics == CN_HOLONOMIC) theta = atan2((double)a1[VY], (double)a1[VX]);
    else theta = (double)a1[TH];
    double fx = cos(theta), fy = sin(theta);
    double dx = (double)a2[PX] - (double)a1[PX], dy = (double)a2[PY] - (double)a1[PY];
    const double nf = norm2(fx, fy);
    fx = fx / nf; fy = fy / nf;
    const double nd = norm2(dx, dy);
    dx = dx / nd; dy = dy / nd;
    double d = fx * dx + fy * dy;
    if (isnan(d)) return 0;             /* coincident agents: nan <= fov/2 is False */
    if (d < -1.0) d = -1.0;
    if (d > 1.0) d = 1.0;
    const double offset = acos(d);
    return fabs(offset) <= fov / 2.0;
}

/* VelocityRectangle (helper.py:199-231) as 4 vertices, shapely affine arithmetic */
static void velocity_rect(const float *a, double q[4][2])
{
    const double vx = a[VX], vy = a[VY], radius = a[RAD];
    const double w = 2.0 * radius * 1.0;
    const double L = 3.0 * sqrt(vx * vx + vy * vy);
    const double heading = atan2(vy, vx);
    const double dth = heading - PI / 2.0;
    const double xos = (double)a[PX] + radius * cos(heading);
    const double yos = (double)a[PY] + radius * sin(heading);
    /* box(-w/2,-L/2,w/2,L/2), ccw from (maxx,miny) */
    double p[4][2] = {{w / 2, -L / 2}, {w / 2, L / 2}, {-w / 2, L / 2}, {-w / 2, -L / 2}};
    double c = cos(dth), s = sin(dth);
    if (fabs(c) < 2.5e-16) c = 0.0;
    if (fabs(s) < 2.5e-16) s = 0.0;
    for (int k = 0; k < 4; ++k) {
        const double x = p[k][0] + 0.0, y = p[k][1] + L / 2;      /* translate(0, L/2) */
        const double xr = c * x + (-s) * y + 0.0;                 /* rotate about (0,0) */
        const double yr = s * x + c * y + 0.0;
        q[k][0] = xr + xos;                                       /* translate(x_os, y_os) */
        q[k][1] = yr + yos;
    }
}

static int rects_intersect(double a[4][2], double b[4][2])
{
    for (int which = 0; which < 2; ++which) {
        double (*p)[2] = which ? b : a;
        for (int k = 0; k < 4; ++k) {
            const double ax = -(p[(k + 1) & 3][1] - p[k][1]);
            const double ay = p[(k + 1) & 3][0] - p[k][0];
            double a0 = INFINITY, a1 = -INFINITY, b0 = INFINITY, b1 = -INFINITY;
            for (int m = 0; m < 4; ++m) {
                const double va = a[m][0] * ax + a[m][1] * ay;
                const double vb = b[m][0] * ax + b[m][1] * ay;
                if (va < a0) a0 = va;
                if (va > a1) a1 = va;
                if (vb < b0) b0 = vb;
                if (vb > b1) b1 = vb;
            }
            if (a1 < b0 || b1 < a0) return 0;
        }
    }
    return 1;
}

/* check_inside_world (helper.py:42-55) with the exact disc: False iff the disc touches a wall segment */
static int inside_world(double px, double py, double r, double t)
{
    const double w[5][2] = {{-t, -t}, {t, -t}, {t, t}, {-t, t}, {-t, -t}};
    for (int k = 0; k < 4; ++k) {
        const double ax = w[k][0], ay = w[k][1], bx = w[k + 1][0], by = w[k + 1][1];
        const double dx = bx - ax, dy = by - ay;
        const double l2 = dx * dx + dy * dy;
        double tt = ((px - ax) * dx + (py - ay) * dy) / l2;
        if (tt < 0.0) tt = 0.0;
        if (tt > 1.0) tt = 1.0;
        const double cx = ax + tt * dx, cy = ay + tt * dy;
        const double d2 = (px - cx) * (px - cx) + (py - cy) * (py - cy);
        if (d2 <= r * r) return 0;
    }
    return 1;
}

static uint64_t step_key(const Env *e);
static uint64_t running_episode_key(const Env *e);

/* what human i is shown of the others (get_human_actions, crowd_sim.py:1121-1157): the other humans in index order, then the
 * robot when it is visible to humans; anyone outside the human's FOV -- or, for human 0 under humans.random_unobservability,
 * missed with unobservable_chance (one draw per step and neighbour slot) -- is replaced by the dummy agent parked at (7, 7) */
static size_t observed_others(const Env *e, int i, orc_v2 *o_pos, orc_v2 *o_vel, double *o_raw_r)
{
    const CnConfig *cfg = e->cfg;
    const int H = e->H;
    const float *self = e->hum + 9 * i;
    const int limited = cfg->human_fov < 2.0 * PI;
    size_t m = 0;
    for (int j = 0; j <= H; ++j) {
        if (j == i) continue;
        if (j == H && !cfg->robot_visible) break;
        const float *o = (j == H) ? e->rob : e->hum + 9 * j;
        int visible = !limited || detect_visible(cfg, self, o, cfg->human_fov);
        if (cfg->random_unobservability && i == 0) {
            double u[4];
            philox_u01(step_key(e), (uint32_t)m, 0, (uint32_t)e->ctr[0], RNG_UNOBS, u);
            if (u[0] <= cfg->unobservable_chance) visible = 0;
        }
        if (visible) {
            o_pos[m] = orc_mk(o[PX], o[PY]);
            o_vel[m] = orc_mk(o[VX], o[VY]);
            o_raw_r[m] = (double)o[RAD];
        } else { /* dummy_human / dummy_robot .set(7,7,7,7,0,0,0), crowd_sim.py:161-168 */
            o_pos[m] = orc_mk(7.0f, 7.0f);
            o_vel[m] = orc_mk(0.0f, 0.0f);
            o_raw_r[m] = (j == H) ? cfg->robot_radius : cfg->human_radius;
        }
        ++m;
    }
    return m;
}

/* humans.random_policy_changing (crowd_sim.py:463-473): ORCA or social force per human and episode, equal chance; a pure
 * function of (key of the running episode, human) */
static int human_policy_of(const Env *e, int i)
{
    if (!e->cfg->random_policy_changing) return e->cfg->human_policy;
    uint32_t x[4];
    oracle_philox(running_episode_key(e), 0, (uint32_t)i, 0, RNG_POLICY, x);
    return (int)(x[0] & 1u);
}

/* SOCIAL_FORCE.predict (social_force.py:11-63), Python floats -> double; the action is narrowed to the float32 state */
static void human_social_force(const Env *e, int i, size_t m, const orc_v2 *o_pos, const double *o_raw_r, float out[2])
{
    const CnConfig *cfg = e->cfg;
    const float *self = e->hum + 9 * i;
    const double gdx = (double)self[GX] - (double)self[PX], gdy = (double)self[GY] - (double)self[PY];
    const double gdist = sqrt(gdx * gdx + gdy * gdy);
    const double vp = (double)self[VPREF];
    const double cdx = cfg->sf_KI * ((gdx / gdist) * vp - (double)self[VX]);
    const double cdy = cfg->sf_KI * ((gdy / gdist) * vp - (double)self[VY]);
    double ivx = 0.0, ivy = 0.0;
    for (size_t k = 0; k < m; ++k) {
        const double ddx = (double)self[PX] - (double)o_pos[k].x, ddy = (double)self[PY] - (double)o_pos[k].y;
        const double dist = sqrt(ddx * ddx + ddy * ddy);
        const double f = cfg->sf_A * exp(((double)self[RAD] + o_raw_r[k] - dist) / cfg->sf_B);
        ivx += f * (ddx / dist);
        ivy += f * (ddy / dist);
    }
    const double nvx = (double)self[VX] + (cdx + ivx) * cfg->time_step;
    const double nvy = (double)self[VY] + (cdy + ivy) * cfg->time_step;
    const double nrm = sqrt(nvx * nvx + nvy * nvy);
    if (nrm > vp) { out[0] = (float)(nvx / nrm * vp); out[1] = (float)(nvy / nrm * vp); }
    else { out[0] = (float)nvx; out[1] = (float)nvy; }
}

/* human i's action: ORCA.predict (orca.py:64-139) or SOCIAL_FORCE.predict on what it observes */
static void human_action(const Env *e, int i, float out[2])
{
    const CnConfig *cfg = e->cfg;
    const float *self = e->hum + 9 * i;
    orc_v2 o_pos[ORC_MAX_LINES], o_vel[ORC_MAX_LINES];
    double o_raw_r[ORC_MAX_LINES];
    float o_rad[ORC_MAX_LINES];
    const size_t m = observed_others(e, i, o_pos, o_vel, o_raw_r);
    if (human_policy_of(e, i) == CN_POLICY_SOCIAL_FORCE) { human_social_force(e, i, m, o_pos, o_raw_r, out); return; }
    for (size_t k = 0; k < m; ++k) o_rad[k] = (float)(o_raw_r[k] + 0.01 + (double)cfg->orca_safety_space);
    const double gx = (double)self[GX] - (double)self[PX], gy = (double)self[GY] - (double)self[PY];
    const double speed = norm2(gx, gy);
    double pvx = gx, pvy = gy;
    if (speed > 1.0) { pvx = gx / speed; pvy = gy / speed; }
    const float radius = (float)((double)self[RAD] + 0.01 + (double)cfg->orca_safety_space);
    orc_v2 nv;
    orc_new_velocity(orc_mk(self[PX], self[PY]), orc_mk(self[VX], self[VY]), radius, self[VPREF],
                     orc_mk((float)pvx, (float)pvy), m, o_pos, o_vel, o_rad,
                     cfg->orca_neighbor_dist, m, cfg->orca_time_horizon, (float)cfg->time_step, &nv, NULL);
    out[0] = nv.x; out[1] = nv.y;
}

int oracle_orca_one(const CnConfig *cfg, const float *self9, int n_others, const float *others5, float *out_v)
{
    if (n_others > ORC_MAX_LINES) return -1;
    orc_v2 o_pos[ORC_MAX_LINES], o_vel[ORC_MAX_LINES];
    float o_rad[ORC_MAX_LINES];
    for (int k = 0; k < n_others; ++k) {
        const float *o = others5 + 5 * k;
        o_pos[k] = orc_mk(o[0], o[1]);
        o_vel[k] = orc_mk(o[2], o[3]);
        o_rad[k] = (float)((double)o[4] + 0.01 + (double)cfg->orca_safety_space);
    }
    const double gx = (double)self9[GX] - (double)self9[PX], gy = (double)self9[GY] - (double)self9[PY];
    const double speed = norm2(gx, gy);
    double pvx = gx, pvy = gy;
    if (speed > 1.0) { pvx = gx / speed; pvy = gy / speed; }
    const float radius = (float)((double)self9[RAD] + 0.01 + (double)cfg->orca_safety_space);
    orc_v2 nv;
    orc_new_velocity(orc_mk(self9[PX], self9[PY]), orc_mk(self9[VX], self9[VY]), radius, self9[VPREF],
                     orc_mk((float)pvx, (float)pvy), (size_t)n_others, o_pos, o_vel, o_rad,
                     cfg->orca_neighbor_dist, (size_t)n_others, cfg->orca_time_horizon, (float)cfg->time_step, &nv, NULL);
    out_v[0] = nv.x; out_v[1] = nv.y;
    return 0;
}

/* generate_ob (crowd_sim_dict.py:72-103): FOV mask, belief update, obs dict (float32) */
static void generate_ob(const Env *e, int reset, int idx, const CnObsOut *obs)
{
    const CnConfig *cfg = e->cfg;
    const int H = e->H;
    uint32_t mask = 0;
    for (int i = 0; i < H; ++i) {
        const float *h = e->hum + 9 * i;
        float *b = e->bel + 5 * i;
        if (detect_visible(cfg, e->rob, h, cfg->robot_fov)) {
            mask |= 1u << i;
            b[0] = h[PX]; b[1] = h[PY]; b[2] = h[VX]; b[3] = h[VY]; b[4] = h[RAD];
        } else if (reset) {
            b[0] = 15.0f; b[1] = 15.0f; b[2] = 0.0f; b[3] = 0.0f; b[4] = 0.3f;
        } else {
            b[0] = (float)((double)b[0] + (double)b[2] * cfg->time_step);
            b[1] = (float)((double)b[1] + (double)b[3] * cfg->time_step);
        }
    }
    if (!obs) return;
    if (obs->robot_node) {
        float *o = obs->robot_node + 7 * (size_t)idx;
        o[0] = e->rob[PX]; o[1] = e->rob[PY]; o[2] = e->rob[RAD]; o[3] = e->rob[GX];
        o[4] = e->rob[GY]; o[5] = e->rob[VPREF]; o[6] = e->rob[TH];
    }
    if (obs->temporal_edges) {
        obs->temporal_edges[2 * (size_t)idx + 0] = e->rob[VX];
        obs->temporal_edges[2 * (size_t)idx + 1] = e->rob[VY];
    }
    if (obs->spatial_edges) {
        for (int i = 0; i < H; ++i) {
            float *o = obs->spatial_edges + 2 * ((size_t)idx * H + i);
            o[0] = (float)((double)e->bel[5 * i + 0] - (double)e->rob[PX]);
            o[1] = (float)((double)e->bel[5 * i + 1] - (double)e->rob[PY]);
        }
    }
    if (obs->visible_mask) obs->visible_mask[idx] = mask;
}

/* create_agent_attributes (crowd_sim.py:296-357) from 6 uniforms: u[0..1] noise, u[2..5] scenario draws */
static void agent_attributes(const CnConfig *cfg, int scenario, double h_radius, double h_vpref, double robot_radius,
                             const double u[6], double *px, double *py, double *gx, double *gy, double *heading, double *vpref_out)
{
    double v_pref = (h_vpref == 0.0) ? 1.0 : h_vpref;
    const double nx = (u[0] - 0.5) * v_pref, ny = (u[1] - 0.5) * v_pref;
    const double R = cfg->circle_radius, sw = cfg->square_width;
#define WORLD(uu) (((uu) - 0.5) * sw / 2.0)
    *heading = 0.0;
    *px = *py = *gx = *gy = 0.0;
    switch (scenario) {
    case CN_SCN_CIRCLE_CROSSING: {
        const double angle = u[2] * PI * 2.0;
        *px = R * cos(angle) + nx; *py = R * sin(angle) + ny; *gx = -*px; *gy = -*py;
    } break;
    case CN_SCN_SQUARE_CROSSING:
        *px = WORLD(u[2]) * 0.4 + nx; *py = WORLD(u[3]) * 0.4 + ny;
        *gx = WORLD(u[4]) * 0.4 + nx; *gy = WORLD(u[5]) * 0.4 + ny;
        break;
    case CN_SCN_PARALLEL_TRAFFIC: {
        const double sign = (u[2] >= 0.5) ? 1.0 : -1.0;
        *px = WORLD(u[3]) * 0.4 + nx; *py = sign * (u[4] * 3.0 + 1.0 + ny); *gx = *px; *gy = -*py;
    } break;
    case CN_SCN_PERPENDICULAR_TRAFFIC: {
        const double sign = (u[2] >= 0.5) ? 1.0 : -1.0;
        *px = sign * (u[3] * 3.0 + 1.0 + nx); *gx = -*px; *py = WORLD(u[4]) * 0.4 + ny; *gy = *py;
    } break;
    case CN_SCN_SIDE_PREF_PASSING:
    case CN_SCN_SIDE_PREF_OVERTAKING: {
        const double min_x = -(robot_radius + h_radius), max_x = -min_x;
        const double hx = (max_x - min_x) * u[2] + min_x;
        *px = hx; *gx = hx;
        if (scenario == CN_SCN_SIDE_PREF_PASSING) { *py = R; *gy = -R; *heading = -PI / 2.0; }
        else { *py = -R + 2.0; *gy = R + 2.0; *heading = PI / 2.0; v_pref = 0.3; }
    } break;
    case CN_SCN_SIDE_PREF_CROSSING: {
        const double min_x = -(R + robot_radius + h_radius), max_x = -(R - robot_radius - h_radius);
        const double hx = (max_x - min_x) * u[2] + min_x;
        *px = hx; *gx = -hx; *py = 0.0; *gy = 0.0;
    } break;
    default: break;
    }
#undef WORLD
    *vpref_out = v_pref;
}

static uint64_t episode_key(const Env *e)
{
    return e->cfg->seed_offset + (uint64_t)(uint32_t)e->ctr[2] + e->cfg->base_seed + e->env_gid;
}

/* reset() used episode_key() and then advanced case_counter by nenv: the key of the episode the env is IN is one stride back */
static uint64_t running_episode_key(const Env *e)
{
    const uint64_t cs = e->cfg->case_size;
    const uint64_t prev = ((uint64_t)(uint32_t)e->ctr[2] + cs - (uint64_t)(uint32_t)e->cfg->nenv % cs) % cs;
    return e->cfg->seed_offset + prev + e->cfg->base_seed + e->env_gid;
}

static uint64_t step_key(const Env *e)
{
    return (e->cfg->base_seed + e->env_gid) ^ ((uint64_t)(uint32_t)e->ctr[1] << 32) ^ 0x5EEDC0DE00000000ull;
}

/* check_collision_group (crowd_sim.py:520-538) / check_collision_group_goal (:541-552): a disc of `radius` at (x, y) against
 * the circle groups (<= group radius + radius + margin) and, when `movers`, the first n_h humans that are not obstacles */
static int collides_with_groups(const Env *e, double x, double y, double radius, double margin, int n_h, int movers)
{
    for (int g = 0; g < CN_MAX_GROUPS; ++g) {
        const float *q = e->grp + 4 * g;
        if (q[3] == 0.0f) break;
        if (norm2(x - (double)q[1], y - (double)q[2]) <= (double)q[0] + radius + margin) return 1;
    }
    if (movers)
        for (int k = 0; k < n_h; ++k) {
            const float *a = e->hum + 9 * k;
            if (a[VPREF] == 0.0f) continue;       /* isObstacle */
            if (norm2(x - (double)a[PX], y - (double)a[PY]) <= (double)a[RAD] + radius) return 1;
        }
    return 0;
}

/* does goal (gx,gy) of human i collide with any other agent's position or goal (crowd_sim.py:750-759, 795-804) */
static int goal_collides(const Env *e, int i, double gx, double gy)
{
    if (e->cfg->group_human)                  /* crowd_sim.py:747-748, 792-793: the human itself is not excluded there */
        return collides_with_groups(e, gx, gy, (double)e->hum[9 * i + RAD], 2 * 0.5, e->H, 1);
    const double ri = e->hum[9 * i + RAD];
    const double dd = e->cfg->discomfort_dist;
    {
        const float *a = e->rob;
        const double md = ri + (double)a[RAD] + dd;
        if (norm2(gx - a[PX], gy - a[PY]) < md || norm2(gx - a[GX], gy - a[GY]) < md) return 1;
    }
    for (int k = 0; k < e->H; ++k) {
        if (k == i) continue;
        const float *a = e->hum + 9 * k;
        const double md = ri + (double)a[RAD] + dd;
        if (norm2(gx - a[PX], gy - a[PY]) < md || norm2(gx - a[GX], gy - a[GY]) < md) return 1;
    }
    return 0;
}

/* update_human_goals_randomly + update_human_goal triggers (crowd_sim_dict.py:261-269) */
static uint32_t goal_updates(const Env *e)
{
    const CnConfig *cfg = e->cfg;
    const int H = e->H;
    const uint32_t s = (uint32_t)e->ctr[0];  /* steps taken so far (post-increment) */
    const uint64_t key = step_key(e);
    uint32_t changed = 0;
    double u[4];
    if (cfg->random_goal_changing && s < 32u * CN_STEP_TABLE_WORDS &&
        ((cfg->goal_change_steps[s >> 5] >> (s & 31)) & 1u)) {
        for (int i = 0; i < H; ++i) {
            float *h = e->hum + 9 * i;
            if (h[VPREF] == 0.0f) continue;
            philox_u01(key, RNG_DECISION, (uint32_t)i, s, RNG_GOAL_RANDOM, u);
            if (!(u[0] <= cfg->goal_change_chance)) continue;
            for (int t = 0; t < cfg->max_goal_tries; ++t) {
                philox_u01(key, (uint32_t)t, (uint32_t)i, s, RNG_GOAL_RANDOM, u);
                const double angle = u[0] * PI * 2.0;
                const double v_pref = (h[VPREF] == 0.0f) ? 1.0 : (double)h[VPREF];
                const double gx = cfg->circle_radius * cos(angle) + (u[1] - 0.5) * v_pref;
                const double gy = cfg->circle_radius * sin(angle) + (u[2] - 0.5) * v_pref;
                if (!goal_collides(e, i, gx, gy)) {
                    h[GX] = (float)gx; h[GY] = (float)gy; changed |= 1u << i;
                    break;
                }
            }
        }
    }
    if (cfg->end_goal_changing) {
        for (int i = 0; i < H; ++i) {
            float *h = e->hum + 9 * i;
            if (!(norm2((double)h[GX] - (double)h[PX], (double)h[GY] - (double)h[PY]) < (double)h[RAD])) continue;
            philox_u01(key, RNG_DECISION, (uint32_t)i, s, RNG_GOAL_END, u);
            if (!(u[0] <= cfg->end_goal_change_chance)) continue;
            /* humans.random_radii / random_v_pref (crowd_sim.py:779-786): += np.random.uniform(-0.1, 0.1) */
            if (cfg->random_radii) h[RAD] = (float)((double)h[RAD] + (-0.1 + 0.2 * u[1]));
            if (cfg->random_v_pref) h[VPREF] = (float)((double)h[VPREF] + (-0.1 + 0.2 * u[2]));
            for (int t = 0; t < cfg->max_goal_tries; ++t) {
                double ua[4], ub[4], u6[6], px, py, gx, gy, hd, vp;
                philox_u01(key, (uint32_t)(2 * t), (uint32_t)i, s, RNG_GOAL_END, ua);
                philox_u01(key, (uint32_t)(2 * t + 1), (uint32_t)i, s, RNG_GOAL_END, ub);
                u6[0] = ua[0]; u6[1] = ua[1]; u6[2] = ua[2]; u6[3] = ua[3]; u6[4] = ub[0]; u6[5] = ub[1];
                agent_attributes(cfg, e->ctr[3], (double)h[RAD], (double)h[VPREF], (double)e->rob[RAD], u6,
                                 &px, &py, &gx, &gy, &hd, &vp);
                if (!goal_collides(e, i, gx, gy)) {
                    h[GX] = (float)gx; h[GY] = (float)gy; changed |= 1u << i;
                    break;
                }
            }
        }
    }
    return changed;
}

/* generate_robot_humans for the group environment (crowd_sim.py:559-622) with the Philox contract and bounded tries */
static void reset_group_env(const Env *e, int scenario, uint64_t key)
{
    const CnConfig *cfg = e->cfg;
    const int H = e->H;
    float *rob = e->rob;
    double u[4];
    rob[RAD] = (float)cfg->robot_radius;
    rob[VPREF] = (float)cfg->robot_v_pref;
    rob[VX] = 0.0f; rob[VY] = 0.0f;
    memset(e->grp, 0, sizeof(float) * 4 * CN_MAX_GROUPS);
    int left = H, idx = 0, ng = 0;
    while (left > 0) {
        if (left <= 4) {
            /* the remaining humans walk: generate_circle_crossing_human with the group collision rule (crowd_sim.py:371-372) */
            for (; idx < H; ++idx) {
                float *h = e->hum + 9 * idx;
                double v_pref = cfg->human_v_pref, radius = cfg->human_radius;
                if (cfg->randomize_attributes) {
                    philox_u01(key, 0, (uint32_t)idx, 0, RNG_ATTR, u);
                    v_pref = 0.5 + (1.5 - 0.5) * u[0];
                    radius = 0.3 + (0.5 - 0.3) * u[1];
                }
                const float radius_f = (float)radius;
                double px = 0, py = 0, gx = 0, gy = 0, hd = 0, vp = v_pref;
                for (int t = 0; t < cfg->max_spawn_tries; ++t) {
                    double ua[4], ub[4], u6[6];
                    philox_u01(key, (uint32_t)t, (uint32_t)idx, 0, RNG_SPAWN, ua);
                    philox_u01(key, (uint32_t)t, (uint32_t)idx, 1, RNG_SPAWN, ub);
                    u6[0] = ua[0]; u6[1] = ua[1]; u6[2] = ua[2]; u6[3] = ua[3]; u6[4] = ub[0]; u6[5] = ub[1];
                    agent_attributes(cfg, scenario, (double)radius_f, v_pref, (double)rob[RAD], u6, &px, &py, &gx, &gy, &hd, &vp);
                    if (!collides_with_groups(e, px, py, (double)radius_f, 2 * 0.5, idx, 1)) break;
                }
                h[PX] = (float)px; h[PY] = (float)py; h[GX] = (float)gx; h[GY] = (float)gy;
                h[VX] = 0.0f; h[VY] = 0.0f; h[TH] = (float)hd; h[RAD] = radius_f; h[VPREF] = (float)vp;
            }
            left = 0;
        } else {
            /* a circle of circum_num = randint(4, min(left, 10)) static humans (generate_circle_group_obstacle, :476-518) */
            const int max_rand = left < 10 ? left : 10;
            philox_u01(key, (uint32_t)ng, 0, 0, RNG_GROUP, u);
            int circum = 4 + (int)(u[0] * (double)(max_rand - 4));
            if (circum > max_rand - 1) circum = max_rand - 1;
            const double g_radius = cfg->human_radius * 2.0 * circum / (2.0 * PI);
            double cx = 0.0, cy = 0.0;
            for (int t = 0; t < cfg->max_spawn_tries; ++t) {
                philox_u01(key, (uint32_t)t, (uint32_t)ng, 1, RNG_GROUP, u);
                cx = -3.0 + 6.0 * u[0]; cy = -3.0 + 6.0 * u[1];
                int ok = 1;
                for (int g = 0; g < ng && ok; ++g) {
                    const float *q = e->grp + 4 * g;
                    if (norm2(cx - (double)q[1], cy - (double)q[2]) < g_radius + (double)q[0] + 2.0 * cfg->human_radius) ok = 0;
                }
                if (ok) break;
            }
            float *q = e->grp + 4 * ng;
            q[0] = (float)g_radius; q[1] = (float)cx; q[2] = (float)cy; q[3] = 1.0f;
            const double arc = 2.0 * PI / circum;
            for (int k = 0; k < circum; ++k, ++idx) {
                float *h = e->hum + 9 * idx;
                const double angle = arc * k;
                const double px = (double)q[1] + (double)q[0] * cos(angle), py = (double)q[2] + (double)q[0] * sin(angle);
                h[PX] = (float)px; h[PY] = (float)py; h[GX] = h[PX]; h[GY] = h[PY];
                h[VX] = 0.0f; h[VY] = 0.0f; h[TH] = 0.0f; h[RAD] = (float)cfg->human_radius; h[VPREF] = 0.0f;
            }
            left -= circum; ++ng;
        }
    }
    /* robot on the circle of radius 5.5, its goal on the opposite side; both step by 0.2 rad past the groups (:593-620) */
    philox_u01(key, 0, 0, 2, RNG_GROUP, u);
    const double rand_angle = u[0] * PI * 2.0;
    double inc = 0.0, px = 0.0, py = 0.0, gx = 0.0, gy = 0.0;
    for (int t = 0; t < 64; ++t) {
        px = cos(rand_angle + inc) * 5.5; py = sin(rand_angle + inc) * 5.5;
        if (!collides_with_groups(e, px, py, cfg->robot_radius, 2 * 0.5, H, 1)) break;
        inc = inc + 0.2;
    }
    inc = inc + PI;
    for (int t = 0; t < 64; ++t) {
        gx = cos(rand_angle + inc) * 5.5; gy = sin(rand_angle + inc) * 5.5;
        if (!collides_with_groups(e, gx, gy, cfg->robot_radius, 4 * 0.5, H, 0)) break;
        inc = inc + 0.2;
    }
    rob[PX] = (float)px; rob[PY] = (float)py; rob[GX] = (float)gx; rob[GY] = (float)gy; rob[TH] = (float)(PI / 2.0);
}

/* CrowdSimDict.reset (crowd_sim_dict.py:105-203) with the Philox contract */
static void reset_env(const Env *e, int idx, const CnObsOut *obs)
{
    const CnConfig *cfg = e->cfg;
    const int H = e->H;
    const uint64_t key = episode_key(e);
    double u[4];
    philox_u01(key, 0, 0, 0, RNG_RESET, u);
    int scn_idx;
    if (cfg->social_metrics) scn_idx = (int)((uint32_t)e->ctr[1] % 4u);
    else { scn_idx = (int)(u[0] * cfg->n_scenarios); if (scn_idx >= cfg->n_scenarios) scn_idx = cfg->n_scenarios - 1; }
    const int scenario = cfg->scenarios[scn_idx];
    e->ctr[3] = scenario;
    e->ctr[0] = 0;           /* global_time = 0 */
    e->ext[0] = 0.0f;        /* desiredVelocity = [0, 0] */
    const double R = cfg->circle_radius;
    float *rob = e->rob;
    if (cfg->group_human) { reset_group_env(e, scenario, key); goto spawned; }
    rob[RAD] = (float)cfg->robot_radius;
    rob[VPREF] = (float)cfg->robot_v_pref;
    rob[VX] = 0.0f; rob[VY] = 0.0f;
    if (cfg->kinematics == CN_UNICYCLE) {
        const double angle = u[1] * PI * 2.0;
        const double px = R * cos(angle), py = R * sin(angle);
        double gx = 0.0, gy = 0.0;
        for (int t = 0; t < cfg->max_robot_tries; ++t) {
            double v[4];
            philox_u01(key, (uint32_t)t, 1, 0, RNG_RESET, v);
            gx = -R + 2.0 * R * v[0]; gy = -R + 2.0 * R * v[1];
            if (norm2(px - gx, py - gy) >= 6.0) break;
        }
        rob[PX] = (float)px; rob[PY] = (float)py; rob[GX] = (float)gx; rob[GY] = (float)gy;
        rob[TH] = (float)(u[2] * 2.0 * PI);
    } else if (cfg->social_metrics || cfg->side_preference) {
        rob[PX] = 0.0f; rob[PY] = (float)(-R); rob[GX] = 0.0f; rob[GY] = (float)R; rob[TH] = (float)(PI / 2.0);
    } else {
        double px = 0, py = 0, gx = 0, gy = 0;
        for (int t = 0; t < cfg->max_robot_tries; ++t) {
            double v[4];
            philox_u01(key, (uint32_t)t, 1, 0, RNG_RESET, v);
            px = -R + 2.0 * R * v[0]; py = -R + 2.0 * R * v[1]; gx = -R + 2.0 * R * v[2]; gy = -R + 2.0 * R * v[3];
            if (norm2(px - gx, py - gy) >= 6.0) break;
        }
        rob[PX] = (float)px; rob[PY] = (float)py; rob[GX] = (float)gx; rob[GY] = (float)gy; rob[TH] = (float)(PI / 2.0);
    }
    for (int i = 0; i < H; ++i) {
        float *h = e->hum + 9 * i;
        double v_pref = cfg->human_v_pref, radius = cfg->human_radius;
        if (cfg->randomize_attributes) {
            philox_u01(key, 0, (uint32_t)i, 0, RNG_ATTR, u);
            v_pref = 0.5 + (1.5 - 0.5) * u[0];
            radius = 0.3 + (0.5 - 0.3) * u[1];
        }
        /* the min-distance test reads the float32 radius the state will hold */
        const float radius_f = (float)radius;
        double px = 0, py = 0, gx = 0, gy = 0, hd = 0, vp = v_pref;
        for (int t = 0; t < cfg->max_spawn_tries; ++t) {
            double ua[4], ub[4], u6[6];
            philox_u01(key, (uint32_t)t, (uint32_t)i, 0, RNG_SPAWN, ua);
            philox_u01(key, (uint32_t)t, (uint32_t)i, 1, RNG_SPAWN, ub);
            u6[0] = ua[0]; u6[1] = ua[1]; u6[2] = ua[2]; u6[3] = ua[3]; u6[4] = ub[0]; u6[5] = ub[1];
            agent_attributes(cfg, scenario, (double)radius_f, v_pref, (double)rob[RAD], u6, &px, &py, &gx, &gy, &hd, &vp);
            int collide = 0;
            {
                const double md = (cfg->kinematics == CN_UNICYCLE) ? R / 2.0
                                                                   : (double)radius_f + (double)rob[RAD] + cfg->discomfort_dist;
                if (norm2(px - (double)rob[PX], py - (double)rob[PY]) < md) collide = 1;
            }
            for (int k = 0; k < i && !collide; ++k) {
                const float *a = e->hum + 9 * k;
                const double md = (double)radius_f + (double)a[RAD] + cfg->discomfort_dist;
                if (norm2(px - (double)a[PX], py - (double)a[PY]) < md) collide = 1;
            }
            if (!collide) break;
        }
        h[PX] = (float)px; h[PY] = (float)py; h[GX] = (float)gx; h[GY] = (float)gy;
        h[VX] = 0.0f; h[VY] = 0.0f; h[TH] = (float)hd; h[RAD] = radius_f; h[VPREF] = (float)vp;
    }
spawned:
    /* case_counter[phase] = (case_counter + nenv) % case_size (crowd_sim_dict.py:162-164) */
    e->ctr[2] = (int32_t)(uint32_t)(((uint64_t)(uint32_t)e->ctr[2] + (uint64_t)cfg->nenv) % cfg->case_size);
    generate_ob(e, 1, idx, obs);
    e->ext[1] = (float)(-fabs(norm2((double)rob[PX] - (double)rob[GX], (double)rob[PY] - (double)rob[GY])));
    e->ctr[1] += 1;          /* scenario_counter */
    *e->ep_ret = 0.0f;
}

static void step_env(const Env *e, int idx, const float *action, const CnStepOut *out, int auto_reset)
{
    const CnConfig *cfg = e->cfg;
    const int H = e->H;
    float *rob = e->rob;
    const double dt = cfg->time_step;

    /* clip_action (srnn.py:18-48): float32 arithmetic on the float32 action array */
    float a0 = action[0], a1 = action[1];
    double act_v = 0.0, act_r = 0.0;      /* unicycle ActionRot */
    double avx, avy;                      /* world-frame velocity the robot will have (holonomic action / patch P1) */
    if (cfg->kinematics == CN_HOLONOMIC) {
        const float nrm = sqrtf(a0 * a0 + a1 * a1);
        const float vp = rob[VPREF];
        if (nrm > vp) { a0 = a0 / nrm * vp; a1 = a1 / nrm * vp; }
        avx = a0; avy = a1;
    } else {
        const float lim = 0.1f;
        if (a0 < -lim) a0 = -lim;
        if (a0 > lim) a0 = lim;
        if (a1 < -lim) a1 = -lim;
        if (a1 > lim) a1 = lim;
        double dv = (double)e->ext[0] + (double)a0;
        const double vp = rob[VPREF];
        if (dv < -vp) dv = -vp;
        if (dv > vp) dv = vp;
        e->ext[0] = (float)dv;
        act_v = (double)e->ext[0]; act_r = (double)a1;
        /* declared oracle patch P1 (crowd_sim.py:1004-1005,1023 read .vx/.vy of an ActionRot) */
        avx = act_v * cos((double)rob[TH] + act_r);
        avy = act_v * sin((double)rob[TH] + act_r);
    }

    /* human actions on the pre-step state */
    float hact[CN_MAX_HUMANS][2];
    for (int i = 0; i < H; ++i) human_action(e, i, hact[i]);

    /* calc_reward on the pre-step state (crowd_sim.py:907-1094) */
    double dmin = INFINITY;
    int collision = 0, vec_viol = 0, agg_nav = 0;
    double rvr[4][2];
    velocity_rect(rob, rvr);
    for (int i = 0; i < H; ++i) {
        const float *h = e->hum + 9 * i;
        const double dx = (double)h[PX] - (double)rob[PX], dy = (double)h[PY] - (double)rob[PY];
        const double closest = sqrt(dx * dx + dy * dy) - (double)h[RAD] - (double)rob[RAD];
        if (closest < 0.0) { collision = 1; break; }
        else if (closest < dmin) dmin = closest;
        double hvr[4][2];
        velocity_rect(h, hvr);
        if (rects_intersect(rvr, hvr)) vec_viol += 1;
        if (!(norm2((double)h[PX] - (double)h[GX], (double)h[PY] - (double)h[GY]) < (double)h[RAD])) agg_nav += 1;
    }
    const double dgoal = norm2((double)rob[PX] - (double)rob[GX], (double)rob[PY] - (double)rob[GY]);
    const int reaching_goal = dgoal < (double)rob[RAD];
    if (!reaching_goal) agg_nav += 1;

    /* robot end position (Agent.compute_position, agent.py:172-196) */
    double npx, npy, nth = rob[TH], nvx, nvy;
    if (cfg->kinematics == CN_HOLONOMIC) {
        npx = (double)rob[PX] + avx * dt; npy = (double)rob[PY] + avy * dt; nvx = avx; nvy = avy;
    } else {
        double Rr;
        if (fabs(act_r) < 0.0001) Rr = 0.0;
        else { const double w = act_r / dt; Rr = act_v / w; }
        const double th = rob[TH];
        npx = (double)rob[PX] - Rr * sin(th) + Rr * sin(th + act_r);
        npy = (double)rob[PY] + Rr * cos(th) - Rr * cos(th + act_r);
        nth = fmod(th + act_r, 2.0 * PI);
        if (nth < 0.0) nth += 2.0 * PI;     /* Python % */
        nvx = act_v * cos(nth); nvy = act_v * sin(nth);
    }

    float *info = out->info ? out->info + (size_t)CN_INFO_DIM * idx : NULL;
    float infobuf[CN_INFO_DIM];
    if (!info) info = infobuf;
    memset(info, 0, sizeof(float) * CN_INFO_DIM);
    if (cfg->side_preference) {
        const float *h = e->hum;
        if (npy <= (double)h[PY] + (double)h[RAD] && npy >= (double)h[PY] - (double)h[RAD]) {
            if (npx < (double)h[PX]) info[CN_INFO_SIDE_LEFT] = 1.0f; else info[CN_INFO_SIDE_RIGHT] = 1.0f;
        }
        info[CN_INFO_SEPARATION] = (float)norm2((double)h[PX] - (double)rob[PX], (double)h[PY] - (double)rob[PY]);
    }
    info[CN_INFO_DMIN] = (float)dmin;
    info[CN_INFO_AGGREGATE_NAV_TIME] = (float)agg_nav;
    info[CN_INFO_PATH_VIOLATION] = (float)vec_viol;
    info[CN_INFO_PERSONAL_VIOLATION] = (dmin < cfg->min_personal_space) ? 1.0f : 0.0f;
    {
        const double ax = avx - (double)rob[VX], ay = avy - (double)rob[VY];
        const double dax = ax - (double)e->ext[2], day = ay - (double)e->ext[3];
        info[CN_INFO_JERK_COST] = (float)(dax * dax + day * day);
        e->ext[2] = (float)ax; e->ext[3] = (float)ay;
    }
    info[CN_INFO_DIST_TO_GOAL] = (float)dgoal;
    const int inside = inside_world((double)rob[PX], (double)rob[PY], (double)rob[RAD], cfg->square_width / 2.0);
    info[CN_INFO_SPEED_VIOLATION] = (sqrt(avx * avx + avy * avy) > cfg->max_walking_speed) ? 1.0f : 0.0f;

    double reward;
    int done, event;
    const int s = e->ctr[0];
    if (s >= cfg->timeout_step) { reward = 0.0; done = 1; event = CN_EV_TIMEOUT; }
    else if (collision || !inside) { reward = cfg->collision_penalty; done = 1; event = CN_EV_COLLISION; }
    else if (reaching_goal) {
        reward = cfg->success_reward;
        if (cfg->time_factor) reward *= (cfg->time_limit - (double)s * dt) / cfg->time_limit;
        done = 1; event = CN_EV_REACH_GOAL;
    } else if (dmin < cfg->discomfort_dist) {
        reward = (dmin - cfg->discomfort_dist) * cfg->discomfort_penalty_factor; done = 0; event = CN_EV_DANGER;
    } else {
        reward = 0.0;
        if (cfg->potential_based) {
            reward = cfg->potential_factor * (-fabs(dgoal) - (double)e->ext[1]);
            e->ext[1] = (float)(-fabs(dgoal));
        } else if (cfg->exponential) {
            reward = cfg->exp_factor * (1.0 - pow(dgoal / cfg->exp_denom, 0.4));
        }
        done = 0; event = CN_EV_NOTHING;
    }
    if (cfg->kinematics == CN_UNICYCLE) {
        const double r_spin = -2.0 * act_r * act_r;
        const double r_back = (act_v < 0.0) ? -2.0 * fabs(act_v) : 0.0;
        reward = reward + r_spin + r_back;
    }

    /* apply actions (agent.py:198-212) */
    rob[PX] = (float)npx; rob[PY] = (float)npy; rob[VX] = (float)nvx; rob[VY] = (float)nvy; rob[TH] = (float)nth;
    for (int i = 0; i < H; ++i) {
        float *h = e->hum + 9 * i;
        h[PX] = (float)((double)h[PX] + (double)hact[i][0] * dt);
        h[PY] = (float)((double)h[PY] + (double)hact[i][1] * dt);
        h[VX] = hact[i][0]; h[VY] = hact[i][1];
    }
    e->ctr[0] = s + 1;

    generate_ob(e, 0, idx, &out->obs);
    const uint32_t changed = goal_updates(e);
    if (out->goal_changed) out->goal_changed[idx] = changed;

    *e->ep_ret = (float)((double)*e->ep_ret + reward);
    if (out->reward) out->reward[idx] = (float)reward;
    if (out->done) out->done[idx] = (uint8_t)done;
    if (out->event) out->event[idx] = event;
    if (out->scenario) out->scenario[idx] = e->ctr[3];
    if (out->episode_return) out->episode_return[idx] = *e->ep_ret;
    if (out->episode_length) out->episode_length[idx] = e->ctr[0];
    if (done && auto_reset) reset_env(e, idx, &out->obs);
}

static Env make_env(const CnConfig *cfg, const OrStateView *st, int idx)
{
    Env e;
    e.cfg = cfg;
    e.H = cfg->human_num;
    e.rob = st->robot + 9 * (size_t)idx;
    e.hum = st->humans + 9 * (size_t)idx * e.H;
    e.bel = st->belief + 5 * (size_t)idx * e.H;
    e.ext = st->extras + 4 * (size_t)idx;
    e.ctr = st->counters + 4 * (size_t)idx;
    e.ep_ret = st->episode_return + idx;
    e.grp = st->groups ? st->groups + 4 * (size_t)CN_MAX_GROUPS * idx : NULL;
    e.env_gid = (uint64_t)(uint32_t)(cfg->env_id_offset + idx);
    return e;
}

static int check_cfg(const CnConfig *cfg)
{
    if (!cfg || cfg->abi_version != CN_ABI_VERSION) return -1;
    if (cfg->human_num < 1 || cfg->human_num > CN_MAX_HUMANS) return -1;
    if (cfg->human_num + (cfg->robot_visible ? 1 : 0) > CN_MAX_HUMANS) return -1;
    return 0;
}

/* envs are independent: split the index range over plain pthreads */
typedef struct {
    const CnConfig *cfg; const OrStateView *st; const float *action; const CnStepOut *out;
    const uint8_t *mask; const CnObsOut *obs; int auto_reset, lo, hi, is_reset;
} Job;

static void *job_main(void *arg)
{
    const Job *j = (const Job *)arg;
    for (int idx = j->lo; idx < j->hi; ++idx) {
        Env e = make_env(j->cfg, j->st, idx);
        if (j->is_reset) { if (!j->mask || j->mask[idx]) reset_env(&e, idx, j->obs); }
        else step_env(&e, idx, j->action + 2 * (size_t)idx, j->out, j->auto_reset);
    }
    return NULL;
}

static void run_jobs(Job proto, int n_envs, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads > n_envs) n_threads = n_envs > 0 ? n_envs : 1;
    pthread_t tid[256];
    Job jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = proto;
        jobs[t].lo = (int)((long long)n_envs * t / n_threads);
        jobs[t].hi = (int)((long long)n_envs * (t + 1) / n_threads);
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&tid[t], NULL, job_main, &jobs[t]);
    job_main(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(tid[t], NULL);
}

int oracle_step(const CnConfig *cfg, int n_envs, const OrStateView *st, const float *action,
                const CnStepOut *out, int auto_reset, int n_threads)
{
    if (check_cfg(cfg)) return -1;
    Job j; memset(&j, 0, sizeof(j));
    j.cfg = cfg; j.st = st; j.action = action; j.out = out; j.auto_reset = auto_reset; j.is_reset = 0;
    run_jobs(j, n_envs, n_threads);
    return 0;
}

int oracle_reset(const CnConfig *cfg, int n_envs, const OrStateView *st, const uint8_t *mask,
                 const CnObsOut *obs, int n_threads)
{
    if (check_cfg(cfg)) return -1;
    Job j; memset(&j, 0, sizeof(j));
    j.cfg = cfg; j.st = st; j.mask = mask; j.obs = obs; j.is_reset = 1;
    run_jobs(j, n_envs, n_threads);
    return 0;
}

int oracle_observe(const CnConfig *cfg, int n_envs, const OrStateView *st, const CnObsOut *obs, int n_threads)
{
    if (check_cfg(cfg)) return -1;
    (void)n_threads;
    for (int idx = 0; idx < n_envs; ++idx) {
        Env e = make_env(cfg, st, idx);
        generate_ob(&e, 1, idx, obs);
    }
    return 0;
}
