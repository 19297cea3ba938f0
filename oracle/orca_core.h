/*
 * oracle/orca_core.h -- TEST INFRASTRUCTURE (CPU oracle). Not product code.
 *
 * Plain-C fp32 restatement of the ORCA solve that the reference performs
 * through the third-party `rvo2` module (github.com/sybrenstuvel/Python-RVO2,
 * UN-PINNED: cloned at HEAD by /root/reference/setup/full_setup.sh:40-44; it
 * wraps RVO2 Library 2.0.x).  The library source is NOT present in
 * /root/reference, so this file restates the published RVO2 algorithm
 * (Agent::computeNeighbors / Agent::computeNewVelocity / linearProgram1-3)
 * as recorded in SURVEY.md Appendix A.  The reference call site it serves is
 * crowd_nav/policy/orca.py:85-136.
 *
 * PARITY UNPINNED against the real rvo2 wheel: the reference ships no golden
 * vectors for ORCA and the wheel cannot be installed here.  Parity of the CUDA
 * path is against THIS restatement.
 *
 * Everything is `float`, evaluated without FMA contraction (build with
 * -ffp-contract=off) like an x86-64 -O3 build of RVO2 without -march.
 */
#ifndef ORACLE_ORCA_CORE_H
#define ORACLE_ORCA_CORE_H

#include <math.h>
#include <stddef.h>

#define ORC_EPSILON 0.00001f
#define ORC_MAX_LINES 64

typedef struct { float x, y; } orc_v2;
typedef struct { orc_v2 point, direction; } orc_line;

static inline orc_v2 orc_mk(float x, float y) { orc_v2 r; r.x = x; r.y = y; return r; }
static inline orc_v2 orc_add(orc_v2 a, orc_v2 b) { return orc_mk(a.x + b.x, a.y + b.y); }
static inline orc_v2 orc_sub(orc_v2 a, orc_v2 b) { return orc_mk(a.x - b.x, a.y - b.y); }
static inline orc_v2 orc_neg(orc_v2 a) { return orc_mk(-a.x, -a.y); }
static inline orc_v2 orc_scale(float s, orc_v2 a) { return orc_mk(s * a.x, s * a.y); }
static inline float orc_dot(orc_v2 a, orc_v2 b) { return a.x * b.x + a.y * b.y; }
static inline float orc_det(orc_v2 a, orc_v2 b) { return a.x * b.y - a.y * b.x; }
static inline float orc_abssq(orc_v2 a) { return orc_dot(a, a); }
static inline float orc_abs(orc_v2 a) { return sqrtf(orc_dot(a, a)); }
/* Vector2 / float is "multiply by reciprocal" in RVO2's Vector2.h */
static inline orc_v2 orc_div(orc_v2 a, float s) { const float inv = 1.0f / s; return orc_mk(a.x * inv, a.y * inv); }
static inline orc_v2 orc_normalize(orc_v2 a) { return orc_div(a, orc_abs(a)); }
static inline float orc_sqr(float a) { return a * a; }

/* linearProgram1: optimise along line `k` subject to lines[0..k) and the speed disc. */
static int orc_lp1(const orc_line *lines, size_t k, float radius, orc_v2 opt, int direction_opt, orc_v2 *result)
{
    const float dp = orc_dot(lines[k].point, lines[k].direction);
    const float disc = orc_sqr(dp) + orc_sqr(radius) - orc_abssq(lines[k].point);
    if (disc < 0.0f) return 0;
    const float s = sqrtf(disc);
    float t_left = -dp - s;
    float t_right = -dp + s;
    for (size_t i = 0; i < k; ++i) {
        const float den = orc_det(lines[k].direction, lines[i].direction);
        const float num = orc_det(lines[i].direction, orc_sub(lines[k].point, lines[i].point));
        if (fabsf(den) <= ORC_EPSILON) {
            if (num < 0.0f) return 0;
            continue;
        }
        const float t = num / den;
        if (den >= 0.0f) { if (t < t_right) t_right = t; }   /* std::min(tRight, t) */
        else             { if (t_left < t) t_left = t; }     /* std::max(tLeft, t)  */
        if (t_left > t_right) return 0;
    }
    if (direction_opt) {
        if (orc_dot(opt, lines[k].direction) > 0.0f)
            *result = orc_add(lines[k].point, orc_scale(t_right, lines[k].direction));
        else
            *result = orc_add(lines[k].point, orc_scale(t_left, lines[k].direction));
    } else {
        const float t = orc_dot(lines[k].direction, orc_sub(opt, lines[k].point));
        if (t < t_left)       *result = orc_add(lines[k].point, orc_scale(t_left, lines[k].direction));
        else if (t > t_right) *result = orc_add(lines[k].point, orc_scale(t_right, lines[k].direction));
        else                  *result = orc_add(lines[k].point, orc_scale(t, lines[k].direction));
    }
    return 1;
}

/* linearProgram2: incremental 2-D LP; returns index of first failing line or n. */
static size_t orc_lp2(const orc_line *lines, size_t n, float radius, orc_v2 opt, int direction_opt, orc_v2 *result)
{
    if (direction_opt) *result = orc_scale(radius, opt);            /* optVelocity * radius */
    else if (orc_abssq(opt) > orc_sqr(radius)) *result = orc_scale(radius, orc_normalize(opt));
    else *result = opt;
    for (size_t i = 0; i < n; ++i) {
        if (orc_det(lines[i].direction, orc_sub(lines[i].point, *result)) > 0.0f) {
            const orc_v2 tmp = *result;
            if (!orc_lp1(lines, i, radius, opt, direction_opt, result)) {
                *result = tmp;
                return i;
            }
        }
    }
    return n;
}

/* linearProgram3 with numObstLines == 0: minimise the maximum penetration. */
static void orc_lp3(const orc_line *lines, size_t n, size_t begin, float radius, orc_v2 *result)
{
    float distance = 0.0f;
    orc_line proj[ORC_MAX_LINES];
    for (size_t i = begin; i < n; ++i) {
        if (orc_det(lines[i].direction, orc_sub(lines[i].point, *result)) > distance) {
            size_t np = 0;
            for (size_t j = 0; j < i; ++j) {
                orc_line l;
                const float d = orc_det(lines[i].direction, lines[j].direction);
                if (fabsf(d) <= ORC_EPSILON) {
                    if (orc_dot(lines[i].direction, lines[j].direction) > 0.0f) continue;
                    l.point = orc_scale(0.5f, orc_add(lines[i].point, lines[j].point));
                } else {
                    const float t = orc_det(lines[j].direction, orc_sub(lines[i].point, lines[j].point)) / d;
                    l.point = orc_add(lines[i].point, orc_scale(t, lines[i].direction));
                }
                l.direction = orc_normalize(orc_sub(lines[j].direction, lines[i].direction));
                proj[np++] = l;
            }
            const orc_v2 tmp = *result;
            if (orc_lp2(proj, np, radius, orc_mk(-lines[i].direction.y, lines[i].direction.x), 1, result) < np)
                *result = tmp;
            distance = orc_det(lines[i].direction, orc_sub(lines[i].point, *result));
        }
    }
}

/*
 * One agent's computeNeighbors + computeNewVelocity.
 * others are visited in index order (the kd-tree visiting order of the real
 * library only matters for exact distSq ties).
 * Returns the number of ORCA lines; *fail_out (optional) gets lp2's return.
 */
static size_t orc_new_velocity(orc_v2 pos, orc_v2 vel, float radius, float max_speed, orc_v2 pref,
                               size_t n_others, const orc_v2 *o_pos, const orc_v2 *o_vel, const float *o_radius,
                               float neighbor_dist, size_t max_neighbors, float time_horizon, float time_step,
                               orc_v2 *new_vel, size_t *fail_out)
{
    /* computeNeighbors / insertAgentNeighbor */
    float nb_d[ORC_MAX_LINES];
    size_t nb_i[ORC_MAX_LINES];
    size_t nn = 0;
    float range_sq = orc_sqr(neighbor_dist);
    if (max_neighbors > 0) {
        for (size_t k = 0; k < n_others; ++k) {
            const float dist_sq = orc_abssq(orc_sub(pos, o_pos[k]));
            if (dist_sq < range_sq) {
                if (nn < max_neighbors) { nb_d[nn] = dist_sq; nb_i[nn] = k; ++nn; }
                size_t i = nn - 1;
                while (i != 0 && dist_sq < nb_d[i - 1]) { nb_d[i] = nb_d[i - 1]; nb_i[i] = nb_i[i - 1]; --i; }
                nb_d[i] = dist_sq; nb_i[i] = k;
                if (nn == max_neighbors) range_sq = nb_d[nn - 1];
            }
        }
    }

    orc_line lines[ORC_MAX_LINES];
    const float inv_tau = 1.0f / time_horizon;
    for (size_t a = 0; a < nn; ++a) {
        const size_t k = nb_i[a];
        const orc_v2 rel_pos = orc_sub(o_pos[k], pos);
        const orc_v2 rel_vel = orc_sub(vel, o_vel[k]);
        const float dist_sq = orc_abssq(rel_pos);
        const float R = radius + o_radius[k];
        const float R2 = orc_sqr(R);
        orc_line line;
        orc_v2 u;
        if (dist_sq > R2) {
            const orc_v2 w = orc_sub(rel_vel, orc_scale(inv_tau, rel_pos));
            const float w_len_sq = orc_abssq(w);
            const float dp1 = orc_dot(w, rel_pos);
            if (dp1 < 0.0f && orc_sqr(dp1) > R2 * w_len_sq) {
                const float w_len = sqrtf(w_len_sq);
                const orc_v2 unit_w = orc_div(w, w_len);
                line.direction = orc_mk(unit_w.y, -unit_w.x);
                u = orc_scale(R * inv_tau - w_len, unit_w);
            } else {
                const float leg = sqrtf(dist_sq - R2);
                if (orc_det(rel_pos, w) > 0.0f)
                    line.direction = orc_div(orc_mk(rel_pos.x * leg - rel_pos.y * R, rel_pos.x * R + rel_pos.y * leg), dist_sq);
                else
                    line.direction = orc_neg(orc_div(orc_mk(rel_pos.x * leg + rel_pos.y * R, -rel_pos.x * R + rel_pos.y * leg), dist_sq));
                const float dp2 = orc_dot(rel_vel, line.direction);
                u = orc_sub(orc_scale(dp2, line.direction), rel_vel);
            }
        } else {
            const float inv_dt = 1.0f / time_step;
            const orc_v2 w = orc_sub(rel_vel, orc_scale(inv_dt, rel_pos));
            const float w_len = orc_abs(w);
            const orc_v2 unit_w = orc_div(w, w_len);
            line.direction = orc_mk(unit_w.y, -unit_w.x);
            u = orc_scale(R * inv_dt - w_len, unit_w);
        }
        line.point = orc_add(vel, orc_scale(0.5f, u));
        lines[a] = line;
    }

    const size_t fail = orc_lp2(lines, nn, max_speed, pref, 0, new_vel);
    if (fail < nn) orc_lp3(lines, nn, fail, max_speed, new_vel);
    if (fail_out) *fail_out = fail;
    return nn;
}

#endif /* ORACLE_ORCA_CORE_H */
