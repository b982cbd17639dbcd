/*
 * hwy_oracle.c — scalar CPU restatement of the HighwayEnv hot path (straight highway
 * family: highway-v0 / highway-fast-v0).  TEST INFRASTRUCTURE ONLY — see hwy_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).  FMA contraction is
 * OFF on purpose: the reference evaluates every expression with separately rounded
 * numpy/Python float64 operations.  The three places where numpy itself fuses (measured
 * in this container, numpy 2.3.5 + its OpenBLAS: 2-vector np.dot = fma(a1*b1 + (a0*b0)),
 * np.linalg.norm = sqrt of that dot; small matmul is NOT fused) use explicit fma().
 *
 * All paths are relative to /root/reference.
 */
#include "hwy_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define VEH_LENGTH 5.0 /* vehicle/kinematics.py:21 */
#define VEH_WIDTH 2.0  /* vehicle/kinematics.py:23 */
#define MAX_SPEED 40.0 /* vehicle/kinematics.py:27 */
#define MIN_SPEED (-40.0)
#define LANE_VEHICLE_LENGTH 5.0 /* road/lane.py:17 */

/* ControlledVehicle constants, vehicle/controller.py:24-33 */
static const double TAU_ACC = 0.6, TAU_HEADING = 0.2, TAU_LATERAL = 0.6;
#define TAU_PURSUIT (0.5 * TAU_HEADING)
#define KP_A (1 / TAU_ACC)
#define KP_HEADING (1 / TAU_HEADING)
#define KP_LATERAL (1 / TAU_LATERAL)
#define MAX_STEERING_ANGLE (M_PI / 3)

/* ------------------------------------------------------------------ numpy RNG */

/* numpy/random/src/pcg64/pcg64.h: 128-bit LCG, XSL-RR output. */
static const uint64_t PCG_MULT_HI = 0x2360ed051fc65da4ULL, PCG_MULT_LO = 0x4385df649fccf645ULL;

uint64_t orc_pcg64_next64(OrcPcg64 *g) {
    unsigned __int128 s = ((unsigned __int128)g->state_hi << 64) | g->state_lo;
    unsigned __int128 m = ((unsigned __int128)PCG_MULT_HI << 64) | PCG_MULT_LO;
    unsigned __int128 inc = ((unsigned __int128)g->inc_hi << 64) | g->inc_lo;
    s = s * m + inc;
    g->state_hi = (uint64_t)(s >> 64);
    g->state_lo = (uint64_t)s;
    uint64_t x = g->state_hi ^ g->state_lo;
    unsigned rot = (unsigned)(g->state_hi >> 58);
    return (x >> rot) | (x << ((-rot) & 63));
}

uint32_t orc_pcg64_next32(OrcPcg64 *g) { /* pcg64_next32: low half first, high half buffered */
    if (g->has_uint32) {
        g->has_uint32 = 0;
        return g->uinteger;
    }
    uint64_t n = orc_pcg64_next64(g);
    g->has_uint32 = 1;
    g->uinteger = (uint32_t)(n >> 32);
    return (uint32_t)n;
}

double orc_pcg64_double(OrcPcg64 *g) {
    return (double)(orc_pcg64_next64(g) >> 11) * (1.0 / 9007199254740992.0);
}

/* Generator.uniform: distributions.c random_uniform = lower + range * next_double */
double orc_rng_uniform(OrcPcg64 *g, double lo, double hi) {
    double range = hi - lo;
    return lo + range * orc_pcg64_double(g);
}

/* Generator.choice(n) / integers(0, n): random_bounded_uint64 with rng = n-1 <= 2^32-1,
 * Lemire rejection on buffered 32-bit draws; rng == 0 consumes nothing. */
int64_t orc_rng_choice(OrcPcg64 *g, int64_t n) {
    uint32_t rng = (uint32_t)(n - 1);
    if (rng == 0) return 0;
    uint32_t rng_excl = rng + 1;
    uint64_t m = (uint64_t)orc_pcg64_next32(g) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        uint32_t threshold = (0xffffffffu - rng) % rng_excl;
        while (leftover < threshold) {
            m = (uint64_t)orc_pcg64_next32(g) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return (int64_t)(m >> 32);
}

/* ------------------------------------------------------------------ utils.py */

static inline double dot2(double a0, double a1, double b0, double b1) {
    return fma(a1, b1, a0 * b0); /* np.dot on 2-vectors, see file header */
}
static inline double norm2(double a0, double a1) { return sqrt(dot2(a0, a1, a0, a1)); }
static inline double clipd(double x, double lo, double hi) { /* np.clip */
    return fmin(fmax(x, lo), hi);
}

/* utils.py:50-56 */
double orc_not_zero(double x) {
    const double eps = 1e-2;
    if (fabs(x) > eps) return x;
    return x >= 0 ? eps : -eps;
}

/* Python / numpy floored float modulo (npy_divmod) */
static inline double py_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* utils.py:59-60 */
double orc_wrap_to_pi(double x) { return py_mod(x + M_PI, 2 * M_PI) - M_PI; }

/* utils.py:31-33 */
static inline double lmap(double v, double x0, double x1, double y0, double y1) {
    return y0 + (v - x0) * (y1 - y0) / (x1 - x0);
}

/* utils.py:77-95 point_in_rotated_rectangle (rotation by -angle... r = [[c,-s],[s,c]],
 * ru = r.dot(point - center): matrix-vector through BLAS gemv; rounding detail is
 * irrelevant to the inclusive-bound KATs, plain expressions are used). */
static int point_in_rotated_rectangle(double px, double py, double cx, double cy, double length,
                                      double width, double angle) {
    double c = cos(angle), s = sin(angle);
    double dx = px - cx, dy = py - cy;
    double rx = c * dx + (-s) * dy, ry = s * dx + c * dy;
    return (-length / 2 <= rx && rx <= length / 2 && -width / 2 <= ry && ry <= width / 2);
}

/* utils.py:160-174 has_corner_inside with rect_corners(include_midpoints, include_center) :128-157 */
static int has_corner_inside(double c1x, double c1y, double l1, double w1, double a1, double c2x,
                             double c2y, double l2, double w2, double a2) {
    double hl = l1 / 2, hw = w1 / 2;
    const double pts[9][2] = {{-hl, -hw}, {-hl, hw}, {hl, hw},  {hl, -hw}, {0, 0},
                              {-hl, 0},   {hl, 0},   {0, -hw}, {0, hw}};
    double c = cos(a1), s = sin(a1);
    for (int k = 0; k < 9; k++) {
        double px = c * pts[k][0] + (-s) * pts[k][1] + c1x;
        double py = s * pts[k][0] + c * pts[k][1] + c1y;
        if (point_in_rotated_rectangle(px, py, c2x, c2y, l2, w2, a2)) return 1;
    }
    return 0;
}

/* utils.py:115-125 */
int orc_rotated_rectangles_intersect(double c1x, double c1y, double l1, double w1, double a1,
                                     double c2x, double c2y, double l2, double w2, double a2) {
    return has_corner_inside(c1x, c1y, l1, w1, a1, c2x, c2y, l2, w2, a2) ||
           has_corner_inside(c2x, c2y, l2, w2, a2, c1x, c1y, l1, w1, a1);
}

/* utils.py:177-185 */
static void project_polygon(const double p[5][2], double ax, double ay, double *mn, double *mx) {
    double lo = 0, hi = 0;
    for (int k = 0; k < 5; k++) {
        double pr = dot2(p[k][0], p[k][1], ax, ay);
        if (k == 0 || pr < lo) lo = pr;
        if (k == 0 || pr > hi) hi = pr;
    }
    *mn = lo;
    *mx = hi;
}

/* utils.py:188-193 */
static inline double interval_distance(double min_a, double max_a, double min_b, double max_b) {
    return min_a < min_b ? min_b - max_a : min_a - max_b;
}

/* utils.py:196-241 are_polygons_intersecting (SAT with velocity extension) */
void orc_polygons_intersecting(const double a[5][2], const double b[5][2], double dax, double day,
                               double dbx, double dby, int *intersecting_out,
                               int *will_intersect_out, double trans[2]) {
    int intersecting = 1, will_intersect = 1;
    double min_distance = INFINITY;
    double tax = 0, tay = 0;
    /* centre difference a[:-1].mean(axis=0) - b[:-1].mean(axis=0): sequential row sum / 4 */
    double cax = (((a[0][0] + a[1][0]) + a[2][0]) + a[3][0]) / 4.0;
    double cay = (((a[0][1] + a[1][1]) + a[2][1]) + a[3][1]) / 4.0;
    double cbx = (((b[0][0] + b[1][0]) + b[2][0]) + b[3][0]) / 4.0;
    double cby = (((b[0][1] + b[1][1]) + b[2][1]) + b[3][1]) / 4.0;
    double dcx = cax - cbx, dcy = cay - cby;
    for (int poly = 0; poly < 2; poly++) {
        const double(*pg)[2] = poly == 0 ? a : b;
        for (int e = 0; e < 4; e++) {
            double nx = -pg[e + 1][1] + pg[e][1];
            double ny = pg[e + 1][0] - pg[e][0];
            double nn = norm2(nx, ny);
            nx /= nn;
            ny /= nn;
            double min_a, max_a, min_b, max_b;
            project_polygon(a, nx, ny, &min_a, &max_a);
            project_polygon(b, nx, ny, &min_b, &max_b);
            if (interval_distance(min_a, max_a, min_b, max_b) > 0) intersecting = 0;
            double vp = dot2(nx, ny, dax - dbx, day - dby);
            if (vp < 0)
                min_a += vp;
            else
                max_a += vp;
            double distance = interval_distance(min_a, max_a, min_b, max_b);
            if (distance > 0) will_intersect = 0;
            if (!intersecting && !will_intersect) break; /* leaves the inner loop only */
            if (fabs(distance) < min_distance) {
                min_distance = fabs(distance);
                if (dot2(dcx, dcy, nx, ny) > 0) {
                    tax = nx;
                    tay = ny;
                } else {
                    tax = -nx;
                    tay = -ny;
                }
            }
        }
    }
    *intersecting_out = intersecting;
    *will_intersect_out = will_intersect;
    if (will_intersect) {
        trans[0] = min_distance * tax;
        trans[1] = min_distance * tay;
    } else {
        trans[0] = trans[1] = 0.0;
    }
}

/* ------------------------------------------------------------------ lanes */

/* StraightLane as built by RoadNetwork.straight_road_network (road/road.py:291-321):
 * lane l: start (0, 4l), end (length, 4l) => direction (1,0), direction_lateral (-0,1),
 * heading 0 (road/lane.py:183-194). */
typedef struct {
    double sx, sy, dx, dy, lx, ly, heading, length, width, speed_limit;
} Lane;

static void make_lanes(const OrcHighwayCfg *c, Lane *lanes) {
    for (int l = 0; l < c->lanes_count; l++) {
        Lane *L = &lanes[l];
        L->sx = 0.0;
        L->sy = l * c->lane_width;
        double ex = 0.0 + c->lane_length, ey = l * c->lane_width;
        L->heading = atan2(ey - L->sy, ex - L->sx);
        L->length = norm2(ex - L->sx, ey - L->sy);
        L->dx = (ex - L->sx) / L->length;
        L->dy = (ey - L->sy) / L->length;
        L->lx = -L->dy;
        L->ly = L->dx;
        L->width = c->lane_width;
        L->speed_limit = c->speed_limit;
    }
}

/* road/lane.py:205-209 StraightLane.local_coordinates */
static inline void lane_local(const Lane *L, double x, double y, double *s, double *lat) {
    double ddx = x - L->sx, ddy = y - L->sy;
    *s = dot2(ddx, ddy, L->dx, L->dy);
    *lat = dot2(ddx, ddy, L->lx, L->ly);
}
static inline double lane_s(const Lane *L, double x, double y) {
    double s, lat;
    lane_local(L, x, y, &s, &lat);
    return s;
}
/* road/lane.py:192-197 position(longitudinal, lateral) */
static inline void lane_position(const Lane *L, double s, double lat, double *x, double *y) {
    *x = (L->sx + s * L->dx) + lat * L->lx;
    *y = (L->sy + s * L->dy) + lat * L->ly;
}
/* road/lane.py:80-102 on_lane */
static inline int lane_on_lane(const Lane *L, double s, double lat, double margin) {
    return fabs(lat) <= L->width / 2 + margin && -LANE_VEHICLE_LENGTH <= s &&
           s < L->length + LANE_VEHICLE_LENGTH;
}
/* road/lane.py:104-118 is_reachable_from (forbidden is always False on the highway) */
static inline int lane_reachable(const Lane *L, double x, double y) {
    double s, lat;
    lane_local(L, x, y, &s, &lat);
    return fabs(lat) <= 2 * L->width && 0 <= s && s < L->length + LANE_VEHICLE_LENGTH;
}
/* road/lane.py:132-143 distance_with_heading */
static inline double lane_distance_with_heading(const Lane *L, double x, double y, double h) {
    double s, r;
    lane_local(L, x, y, &s, &r);
    double angle = fabs(orc_wrap_to_pi(h - L->heading));
    return fabs(r) + fmax(s - L->length, 0) + fmax(0 - s, 0) + 1.0 * angle;
}
/* road/road.py:55-71 get_closest_lane_index: first minimum in enumeration order */
static int closest_lane(const Lane *lanes, int n, double x, double y, double h) {
    int best = 0;
    double bd = 0;
    for (int l = 0; l < n; l++) {
        double d = lane_distance_with_heading(&lanes[l], x, y, h);
        if (l == 0 || d < bd) {
            bd = d;
            best = l;
        }
    }
    return best;
}

/* ------------------------------------------------------------------ world */

typedef struct {
    const OrcHighwayCfg *c;
    OrcHighwayState *s;
    Lane lanes[ORC_MAX_LANES];
    double *act_steer, *act_accel; /* Vehicle.action, persists between act() calls */
    int V;
} World;

/* vehicle/objects.py:183-198 lane_distance_to on self's own lane */
static inline double lane_distance_to(const World *w, int self, int other) {
    const Lane *L = &w->lanes[w->s->lane[self]];
    return lane_s(L, w->s->x[other], w->s->y[other]) - lane_s(L, w->s->x[self], w->s->y[self]);
}

/* road/road.py:483-547 neighbour_vehicles (same-segment search; connected-lane branch off,
 * abstract.py:124).  Ties: front `<=` (later index wins), rear `>` (earlier wins). */
static void neighbour_vehicles(const World *w, int veh, int lane_idx, int *front, int *rear) {
    const Lane *L = &w->lanes[lane_idx];
    double s = lane_s(L, w->s->x[veh], w->s->y[veh]);
    double s_front = 0, s_rear = 0;
    int v_front = -1, v_rear = -1;
    for (int v = 0; v < w->V; v++) {
        if (v == veh) continue;
        double s_v, lat_v;
        lane_local(L, w->s->x[v], w->s->y[v], &s_v, &lat_v);
        if (!lane_on_lane(L, s_v, lat_v, 1.0)) continue;
        if (s <= s_v && (v_front < 0 || s_v <= s_front)) {
            s_front = s_v;
            v_front = v;
        }
        if (s_v < s && (v_rear < 0 || s_v > s_rear)) {
            s_rear = s_v;
            v_rear = v;
        }
    }
    *front = v_front;
    *rear = v_rear;
}

/* vehicle/behavior.py:192-217 desired_gap(ego, front), projected=True */
static double desired_gap(const World *w, int ego, int front) {
    const OrcHighwayCfg *c = w->c;
    const OrcHighwayState *s = w->s;
    double d0 = c->distance_wanted, tau = c->time_wanted;
    double ab = -c->comfort_acc_max * c->comfort_acc_min;
    double ce = cos(s->heading[ego]), se = sin(s->heading[ego]);
    double cf = cos(s->heading[front]), sf = sin(s->heading[front]);
    double dvx = s->speed[ego] * ce - s->speed[front] * cf;
    double dvy = s->speed[ego] * se - s->speed[front] * sf;
    double dv = dot2(dvx, dvy, ce, se);
    return d0 + s->speed[ego] * tau + s->speed[ego] * dv / (2 * sqrt(ab));
}

/* vehicle/behavior.py:150-190 acceleration(); `self_` supplies DELTA (the caller's
 * parameters are used even when reasoning about another vehicle). */
static double idm_acceleration(const World *w, int self_, int ego, int front) {
    const OrcHighwayCfg *c = w->c;
    const OrcHighwayState *s = w->s;
    if (ego < 0) return 0;
    /* getattr(ego_vehicle, "target_speed", 0): a plain Vehicle has none */
    double ego_target_speed = s->kind[ego] == ORC_KIND_VEHICLE ? 0.0 : s->target_speed[ego];
    ego_target_speed = clipd(ego_target_speed, 0, w->lanes[s->lane[ego]].speed_limit);
    double acceleration =
        c->comfort_acc_max *
        (1 - pow(fmax(s->speed[ego], 0) / fabs(orc_not_zero(ego_target_speed)), s->delta[self_]));
    if (front >= 0) {
        double d = lane_distance_to(w, ego, front);
        double q = desired_gap(w, ego, front) / orc_not_zero(d);
        acceleration -= c->comfort_acc_max * pow(q, 2);
    }
    return acceleration;
}

/* vehicle/controller.py:145-187 steering_control */
static double steering_control(const World *w, int v, int target_lane) {
    const OrcHighwayState *s = w->s;
    const Lane *L = &w->lanes[target_lane];
    double lc_s, lc_lat;
    lane_local(L, s->x[v], s->y[v], &lc_s, &lc_lat);
    double lane_next_coords = lc_s + s->speed[v] * TAU_PURSUIT;
    (void)lane_next_coords;
    double lane_future_heading = L->heading; /* StraightLane.heading_at */
    double lateral_speed_command = -KP_LATERAL * lc_lat;
    double heading_command = asin(clipd(lateral_speed_command / orc_not_zero(s->speed[v]), -1, 1));
    double heading_ref = lane_future_heading + clipd(heading_command, -M_PI / 4, M_PI / 4);
    double heading_rate_command = KP_HEADING * orc_wrap_to_pi(heading_ref - s->heading[v]);
    double slip_angle =
        asin(clipd(VEH_LENGTH / 2 / orc_not_zero(s->speed[v]) * heading_rate_command, -1, 1));
    double steering_angle = atan(2 * tan(slip_angle));
    return clipd(steering_angle, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
}

/* vehicle/controller.py:135-143 follow_road.  On the highway graph = {"0": {"1": lanes}}:
 * next_lane (road/road.py:73-136) hits KeyError on graph["1"] and returns the current
 * index, so the target lane never changes. */
static void follow_road(World *w, int v) {
    (void)w;
    (void)v;
}

/* vehicle/behavior.py:265-324 mobil (route is None on the highway => jerk branch) */
static int mobil(const World *w, int v, int lane_index) {
    const OrcHighwayCfg *c = w->c;
    int new_preceding, new_following;
    neighbour_vehicles(w, v, lane_index, &new_preceding, &new_following);
    double new_following_a = idm_acceleration(w, v, new_following, new_preceding);
    double new_following_pred_a = idm_acceleration(w, v, new_following, v);
    if (new_following_pred_a < -c->lane_change_max_braking_imposed) return 0;
    int old_preceding, old_following;
    neighbour_vehicles(w, v, w->s->lane[v], &old_preceding, &old_following);
    double self_pred_a = idm_acceleration(w, v, v, new_preceding);
    double self_a = idm_acceleration(w, v, v, old_preceding);
    double old_following_a = idm_acceleration(w, v, old_following, v);
    double old_following_pred_a = idm_acceleration(w, v, old_following, old_preceding);
    double jerk = self_pred_a - self_a +
                  c->politeness * (new_following_pred_a - new_following_a + old_following_pred_a -
                                   old_following_a);
    if (jerk < c->lane_change_min_acc_gain) return 0;
    return 1;
}

/* vehicle/behavior.py:219-263 change_lane_policy */
static void change_lane_policy(World *w, int v) {
    const OrcHighwayCfg *c = w->c;
    OrcHighwayState *s = w->s;
    if (s->lane[v] != s->target_lane[v]) {
        /* same road always (single road) */
        for (int o = 0; o < w->V; o++) {
            if (o != v && s->lane[o] != s->target_lane[v] && s->kind[o] != ORC_KIND_VEHICLE &&
                s->target_lane[o] == s->target_lane[v]) {
                double d = lane_distance_to(w, v, o);
                double d_star = desired_gap(w, v, o);
                if (0 < d && d < d_star) {
                    s->target_lane[v] = s->lane[v];
                    break;
                }
            }
        }
        return;
    }
    if (!(c->lane_change_delay < s->timer[v])) return; /* utils.do_every, utils.py:27-28 */
    s->timer[v] = 0;
    /* road/road.py:200-211 side_lanes: id-1 then id+1 */
    int cand[2], nc = 0;
    if (s->lane[v] > 0) cand[nc++] = s->lane[v] - 1;
    if (s->lane[v] < c->lanes_count - 1) cand[nc++] = s->lane[v] + 1;
    for (int k = 0; k < nc; k++) {
        if (!lane_reachable(&w->lanes[cand[k]], s->x[v], s->y[v])) continue;
        if (fabs(s->speed[v]) < 1) continue;
        if (mobil(w, v, cand[k])) s->target_lane[v] = cand[k];
    }
}

/* vehicle/behavior.py:93-137 IDMVehicle.act */
static void idm_act(World *w, int v) {
    const OrcHighwayCfg *c = w->c;
    OrcHighwayState *s = w->s;
    if (s->crashed[v]) return;
    follow_road(w, v);
    change_lane_policy(w, v); /* enable_lane_change is True */
    double steering = steering_control(w, v, s->target_lane[v]);
    steering = clipd(steering, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
    int front, rear;
    neighbour_vehicles(w, v, s->lane[v], &front, &rear);
    double acc = idm_acceleration(w, v, v, front);
    if (s->lane[v] != s->target_lane[v]) {
        neighbour_vehicles(w, v, s->target_lane[v], &front, &rear);
        double tacc = idm_acceleration(w, v, v, front);
        acc = fmin(acc, tacc);
    }
    acc = clipd(acc, -c->acc_max, c->acc_max);
    w->act_steer[v] = steering;
    w->act_accel[v] = acc;
}

/* vehicle/controller.py:89-133 ControlledVehicle.act(action); label: 0 LANE_LEFT, 1 IDLE,
 * 2 LANE_RIGHT, -1 None (FASTER/SLOWER are handled by MDPVehicle.act before). */
static void controlled_act(World *w, int v, int label) {
    const OrcHighwayCfg *c = w->c;
    OrcHighwayState *s = w->s;
    follow_road(w, v);
    if (label == 2 || label == 0) {
        int id = s->target_lane[v] + (label == 2 ? 1 : -1);
        if (id < 0) id = 0;
        if (id > c->lanes_count - 1) id = c->lanes_count - 1;
        if (lane_reachable(&w->lanes[id], s->x[v], s->y[v])) s->target_lane[v] = id;
    }
    double steering = steering_control(w, v, s->target_lane[v]);
    double acc = KP_A * (s->target_speed[v] - s->speed[v]); /* speed_control :189-198 */
    steering = clipd(steering, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
    w->act_steer[v] = steering;
    w->act_accel[v] = acc;
}

/* vehicle/controller.py:326-344 speed_to_index (np.round = half to even) */
static int speed_to_index(const OrcHighwayCfg *c, double speed) {
    int n = c->n_target_speeds;
    double x = (speed - c->target_speeds[0]) / (c->target_speeds[n - 1] - c->target_speeds[0]);
    return (int)clipd(rint(x * (n - 1)), 0, n - 1);
}

/* vehicle/controller.py:295-315 MDPVehicle.act(label); envs/common/action.py:204-210 labels */
static void mdp_act(World *w, int v, int action) {
    const OrcHighwayCfg *c = w->c;
    OrcHighwayState *s = w->s;
    if (action == 3 || action == 4) {
        int idx = speed_to_index(c, s->speed[v]) + (action == 3 ? 1 : -1);
        if (idx < 0) idx = 0;
        if (idx > c->n_target_speeds - 1) idx = c->n_target_speeds - 1;
        s->speed_index[0] = idx;
        s->target_speed[v] = c->target_speeds[idx];
        controlled_act(w, v, -1);
    } else {
        controlled_act(w, v, action);
    }
}

/* envs/common/action.py:136-162 ContinuousAction.get_action/act.  The Box action is
 * float32 and NEP-50 keeps lmap (utils.py:31-33) in float32 arithmetic; clip_actions then
 * widens with float(). */
static void continuous_act(World *w, int v, const float *a) {
    const OrcHighwayCfg *c = w->c;
    float a0 = a[0], a1 = a[1];
    if (c->act_clip) {
        a0 = fminf(fmaxf(a0, -1.0f), 1.0f);
        a1 = fminf(fmaxf(a1, -1.0f), 1.0f);
    }
    /* y[0] + (v - x[0]) * (y[1] - y[0]) / (x[1] - x[0]); python scalars are weak */
    float acc = (float)c->acc_lo + (a0 - (-1.0f)) * (float)(c->acc_hi - c->acc_lo) / 2.0f;
    float steer = (float)c->steer_lo + (a1 - (-1.0f)) * (float)(c->steer_hi - c->steer_lo) / 2.0f;
    w->act_accel[v] = (double)acc;
    w->act_steer[v] = (double)steer;
}

/* road/road.py:464-467 Road.act */
static void road_act(World *w) {
    for (int v = 0; v < w->V; v++) {
        switch (w->s->kind[v]) {
        case ORC_KIND_IDM: idm_act(w, v); break;
        case ORC_KIND_MDP: controlled_act(w, v, -1); break; /* MDPVehicle.act(None) */
        default: break;                                     /* Vehicle.act(None): keep action */
        }
    }
}

/* vehicle/kinematics.py:130-177 Vehicle.step (+ IDMVehicle.step timer, behavior.py:139-148) */
static void vehicle_step(World *w, int v, double dt) {
    OrcHighwayState *s = w->s;
    if (s->kind[v] == ORC_KIND_IDM) s->timer[v] += dt;
    /* clip_actions :155-168 */
    if (s->crashed[v]) {
        w->act_steer[v] = 0;
        w->act_accel[v] = -1.0 * s->speed[v];
    }
    if (s->speed[v] > MAX_SPEED)
        w->act_accel[v] = fmin(w->act_accel[v], 1.0 * (MAX_SPEED - s->speed[v]));
    else if (s->speed[v] < MIN_SPEED)
        w->act_accel[v] = fmax(w->act_accel[v], 1.0 * (MIN_SPEED - s->speed[v]));
    double delta_f = w->act_steer[v];
    double beta = atan(1.0 / 2 * tan(delta_f));
    double vx = s->speed[v] * cos(s->heading[v] + beta);
    double vy = s->speed[v] * sin(s->heading[v] + beta);
    s->x[v] += vx * dt;
    s->y[v] += vy * dt;
    if (s->has_impact[v]) {
        s->x[v] += s->impact_x[v];
        s->y[v] += s->impact_y[v];
        s->crashed[v] = 1;
        s->has_impact[v] = 0;
    }
    s->heading[v] += s->speed[v] * sin(beta) / (VEH_LENGTH / 2) * dt;
    s->speed[v] += w->act_accel[v] * dt;
    s->lane[v] = closest_lane(w->lanes, w->c->lanes_count, s->x[v], s->y[v], s->heading[v]);
    /* schema convention: a plain Vehicle has no target lane; mirror lane_index */
    if (s->kind[v] == ORC_KIND_VEHICLE) s->target_lane[v] = s->lane[v];
}

/* vehicle/objects.py:169-181 polygon */
static void polygon(const OrcHighwayState *s, int v, double p[5][2]) {
    static const double loc[4][2] = {{-VEH_LENGTH / 2, -VEH_WIDTH / 2},
                                     {-VEH_LENGTH / 2, +VEH_WIDTH / 2},
                                     {+VEH_LENGTH / 2, +VEH_WIDTH / 2},
                                     {+VEH_LENGTH / 2, -VEH_WIDTH / 2}};
    double c = cos(s->heading[v]), sn = sin(s->heading[v]);
    for (int k = 0; k < 4; k++) {
        p[k][0] = (c * loc[k][0] + (-sn) * loc[k][1]) + s->x[v];
        p[k][1] = (sn * loc[k][0] + c * loc[k][1]) + s->y[v];
    }
    p[4][0] = p[0][0];
    p[4][1] = p[0][1];
}

/* vehicle/objects.py:92-138 handle_collisions + _is_colliding */
static void handle_collisions(World *w, int a, int b, double dt) {
    OrcHighwayState *s = w->s;
    if (!(s->check_collisions[a] || s->check_collisions[b])) return;
    double diag = sqrt(VEH_LENGTH * VEH_LENGTH + VEH_WIDTH * VEH_WIDTH);
    double dist = norm2(s->x[b] - s->x[a], s->y[b] - s->y[a]);
    if (dist > (diag + diag) / 2 + s->speed[a] * dt) return;
    double pa[5][2], pb[5][2], tr[2];
    polygon(s, a, pa);
    polygon(s, b, pb);
    double ca = cos(s->heading[a]), sa = sin(s->heading[a]);
    double cb = cos(s->heading[b]), sb = sin(s->heading[b]);
    int inter, will;
    orc_polygons_intersecting(pa, pb, s->speed[a] * ca * dt, s->speed[a] * sa * dt,
                              s->speed[b] * cb * dt, s->speed[b] * sb * dt, &inter, &will, tr);
    if (will) {
        s->impact_x[a] = tr[0] / 2;
        s->impact_y[a] = tr[1] / 2;
        s->has_impact[a] = 1;
        s->impact_x[b] = -tr[0] / 2;
        s->impact_y[b] = -tr[1] / 2;
        s->has_impact[b] = 1;
    }
    if (inter) {
        s->crashed[a] = 1;
        s->crashed[b] = 1;
    }
}

/* road/road.py:469-481 Road.step */
static void road_step(World *w, double dt) {
    for (int v = 0; v < w->V; v++) vehicle_step(w, v, dt);
    for (int i = 0; i < w->V; i++)
        for (int j = i + 1; j < w->V; j++) handle_collisions(w, i, j, dt);
}

static void world_init(World *w, const OrcHighwayCfg *c, OrcHighwayState *s, double *act_buf) {
    w->c = c;
    w->s = s;
    w->V = c->n_vehicles;
    make_lanes(c, w->lanes);
    w->act_steer = act_buf;
    w->act_accel = act_buf + c->n_vehicles;
    memset(act_buf, 0, sizeof(double) * 2 * c->n_vehicles); /* kinematics.py:44 */
}

int orc_highway_obs_columns(const OrcHighwayCfg *c) { return c->obs_n_features > 0 ? c->obs_n_features : 5; }

/* vehicle/kinematics.py:237-261 Vehicle.to_dict(origin_vehicle, observe_intentions=False) restricted to
 * the configured columns, then normalize_obs (observation.py:207-232) with the configured ranges */
static void observe_row_features(const World *w, int v, int origin, float *out) {
    const OrcHighwayCfg *c = w->c;
    const OrcHighwayState *s = w->s;
    double ch = cos(s->heading[v]), sh = sin(s->heading[v]);
    double x = s->x[v], y = s->y[v], vx = s->speed[v] * ch, vy = s->speed[v] * sh;
    if (origin >= 0) { /* :257-260: only x, y, vx, vy are made relative */
        double co = cos(s->heading[origin]), so = sin(s->heading[origin]);
        x -= s->x[origin];
        y -= s->y[origin];
        vx -= s->speed[origin] * co;
        vy -= s->speed[origin] * so;
    }
    const Lane *L = &w->lanes[s->lane[v]];
    double lon, lat;
    lane_local(L, s->x[v], s->y[v], &lon, &lat); /* lane_offset :228-235 */
    for (int col = 0; col < c->obs_n_features; col++) {
        double val = 0;
        switch (c->obs_feature[col]) {
            case 0: val = 1; break;
            case 1: val = x; break;
            case 2: val = y; break;
            case 3: val = vx; break;
            case 4: val = vy; break;
            case 5: val = s->heading[v]; break;
            case 6: val = ch; break;
            case 7: val = sh; break;
            case 10: val = lon; break;
            case 11: val = lat; break;
            case 12: val = orc_wrap_to_pi(s->heading[v] - L->heading); break; /* lane.py:145-147 */
            default: val = 0; break; /* cos_d, sin_d: no route => destination == position (kinematics.py:205-226) */
        }
        if (c->obs_normalize && c->obs_feature_ranged[col]) {
            val = lmap(val, c->obs_feature_lo[col], c->obs_feature_hi[col], -1, 1);
            if (c->obs_clip) val = clipd(val, -1, 1);
        }
        out[col] = (float)val;
    }
}

/* envs/common/observation.py:234-276 KinematicObservation.observe (features presence,x,y,vx,vy;
 * order "sorted"); road/road.py:421-450 close_objects_to; kinematics.py:237-261 to_dict */
void orc_highway_observe(const OrcHighwayCfg *c, const OrcHighwayState *s, float *obs) {
    World w;
    w.c = c;
    w.s = (OrcHighwayState *)s;
    w.V = c->n_vehicles;
    make_lanes(c, w.lanes);
    int K = c->obs_vehicles_count;
    int V = c->n_vehicles;
    double *rows = (double *)calloc((size_t)K * 5, sizeof(double));
    const int ego = 0;
    double ce = cos(s->heading[ego]), se = sin(s->heading[ego]);
    double evx = s->speed[ego] * ce, evy = s->speed[ego] * se;
    rows[0] = 1;
    rows[1] = s->x[ego];
    rows[2] = s->y[ego];
    rows[3] = evx;
    rows[4] = evy;
    /* candidates in list order, then stable sort by |lane_distance_to| */
    int *cand = (int *)malloc(sizeof(int) * (V > 0 ? V : 1));
    double *key = (double *)malloc(sizeof(double) * (V > 0 ? V : 1));
    int nc = 0;
    for (int v = 0; v < V; v++) {
        if (!(norm2(s->x[v] - s->x[ego], s->y[v] - s->y[ego]) < c->perception_distance)) continue;
        if (v == ego) continue;
        double d = lane_distance_to(&w, ego, v);
        if (!(c->obs_see_behind || -2 * VEH_LENGTH < d)) continue;
        cand[nc] = v;
        key[nc] = fabs(d);
        nc++;
    }
    for (int i = 1; i < nc; i++) { /* insertion sort: stable */
        int cv = cand[i];
        double ck = key[i];
        int j = i - 1;
        while (j >= 0 && key[j] > ck) {
            cand[j + 1] = cand[j];
            key[j + 1] = key[j];
            j--;
        }
        cand[j + 1] = cv;
        key[j + 1] = ck;
    }
    if (c->obs_n_features > 0) { /* configured feature list */
        const int NF = c->obs_n_features;
        for (int k = 0; k < K * NF; k++) obs[k] = 0.0f;
        observe_row_features(&w, ego, -1, obs);
        for (int k = 0; k < nc && k < K - 1; k++)
            observe_row_features(&w, cand[k], c->obs_absolute ? -1 : ego, obs + NF * (k + 1));
        free(rows);
        free(cand);
        free(key);
        return;
    }
    int n_rows = 1;
    for (int k = 0; k < nc && k < K - 1; k++) {
        int v = cand[k];
        double cv = cos(s->heading[v]), sv = sin(s->heading[v]);
        double *r = rows + 5 * n_rows;
        r[0] = 1;
        r[1] = s->x[v];
        r[2] = s->y[v];
        r[3] = s->speed[v] * cv;
        r[4] = s->speed[v] * sv;
        if (!c->obs_absolute) {
            r[1] -= s->x[ego];
            r[2] -= s->y[ego];
            r[3] -= evx;
            r[4] -= evy;
        }
        n_rows++;
    }
    if (c->obs_normalize) { /* normalize_obs :207-232; side lanes of the observer's road */
        double xr = 5.0 * MAX_SPEED, yr = 4.0 /*AbstractLane.DEFAULT_WIDTH*/ * c->lanes_count,
               vr = 2 * MAX_SPEED;
        for (int k = 0; k < n_rows; k++) {
            double *r = rows + 5 * k;
            r[1] = lmap(r[1], -xr, xr, -1, 1);
            r[2] = lmap(r[2], -yr, yr, -1, 1);
            r[3] = lmap(r[3], -vr, vr, -1, 1);
            r[4] = lmap(r[4], -vr, vr, -1, 1);
            if (c->obs_clip)
                for (int f = 1; f < 5; f++) r[f] = clipd(r[f], -1, 1);
        }
    }
    for (int k = 0; k < K * 5; k++) obs[k] = (float)rows[k];
    free(rows);
    free(cand);
    free(key);
}

/* envs/highway_env.py:100-151 _reward/_rewards/_is_terminated/_is_truncated */
static void reward_done(const World *w, double *reward, int32_t *terminated, int32_t *truncated) {
    const OrcHighwayCfg *c = w->c;
    const OrcHighwayState *s = w->s;
    const int ego = 0;
    int lane = s->kind[ego] == ORC_KIND_VEHICLE ? s->lane[ego] : s->target_lane[ego];
    double forward_speed = s->speed[ego] * cos(s->heading[ego]);
    double scaled_speed = lmap(forward_speed, c->reward_speed_lo, c->reward_speed_hi, 0, 1);
    double es, elat;
    lane_local(&w->lanes[s->lane[ego]], s->x[ego], s->y[ego], &es, &elat);
    int on_road = lane_on_lane(&w->lanes[s->lane[ego]], es, elat, 0.0);
    double r_col = (double)(s->crashed[ego] != 0);
    int nl1 = c->lanes_count - 1 > 1 ? c->lanes_count - 1 : 1;
    double r_lane = (double)lane / (double)nl1;
    double r_speed = clipd(scaled_speed, 0, 1);
    double r_road = (double)on_road;
    double reward_ = 0;
    reward_ = reward_ + c->collision_reward * r_col;
    reward_ = reward_ + c->right_lane_reward * r_lane;
    reward_ = reward_ + c->high_speed_reward * r_speed;
    reward_ = reward_ + 0 * r_road; /* config.get("on_road_reward", 0) */
    if (c->normalize_reward)
        reward_ = lmap(reward_, c->collision_reward, c->high_speed_reward + c->right_lane_reward, 0, 1);
    reward_ *= r_road;
    *reward = reward_;
    *terminated = s->crashed[ego] || (c->offroad_terminal && !on_road);
    *truncated = s->time[0] >= c->duration;
}

/* envs/common/abstract.py:259-317 step + _simulate */
void orc_highway_step(const OrcHighwayCfg *c, OrcHighwayState *s, int action_i,
                          const float *action_f, float *obs, double *reward, int32_t *terminated,
                          int32_t *truncated) {
    World w;
    double *act_buf = (double *)malloc(sizeof(double) * 2 * c->n_vehicles);
    world_init(&w, c, s, act_buf);
    int frames = c->simulation_frequency / c->policy_frequency;
    double dt = 1.0 / c->simulation_frequency;
    s->time[0] += 1.0 / c->policy_frequency;
    for (int frame = 0; frame < frames; frame++) {
        if (frame == 0) { /* steps % frames == 0 */
            if (c->action_type == 0)
                mdp_act(&w, 0, action_i);
            else
                continuous_act(&w, 0, action_f);
        }
        road_act(&w);
        road_step(&w, dt);
    }
    if (obs) orc_highway_observe(c, s, obs);
    reward_done(&w, reward, terminated, truncated);
    free(act_buf);
}

void orc_highway_substeps(const OrcHighwayCfg *c, OrcHighwayState *s, int substeps) {
    World w;
    double *act_buf = (double *)malloc(sizeof(double) * 2 * c->n_vehicles);
    world_init(&w, c, s, act_buf);
    double dt = 1.0 / c->simulation_frequency;
    for (int k = 0; k < substeps; k++) {
        road_act(&w);
        road_step(&w, dt);
    }
    free(act_buf);
}

/* envs/highway_env.py:55-98,177-182 _reset; vehicle/kinematics.py:50-104 create_random;
 * vehicle/behavior.py:64-69; vehicle/controller.py:45-48,284-293 */
void orc_highway_reset(const OrcHighwayCfg *c, OrcPcg64 *rng, OrcHighwayState *s) {
    Lane lanes[ORC_MAX_LANES];
    make_lanes(c, lanes);
    int V = c->n_vehicles;
    s->time[0] = 0;
    for (int v = 0; v < V; v++) {
        int is_ego = v == 0;
        /* choice(list(graph.keys())) and choice(list(graph[_from].keys())): one element
         * each => no draw */
        (void)orc_rng_choice(rng, 1);
        (void)orc_rng_choice(rng, 1);
        int id;
        if (is_ego && c->initial_lane_id >= 0)
            id = c->initial_lane_id;
        else
            id = (int)orc_rng_choice(rng, c->lanes_count);
        const Lane *L = &lanes[id];
        double speed;
        if (is_ego)
            speed = c->ego_speed;
        else
            speed = orc_rng_uniform(rng, 0.7 * L->speed_limit, 0.8 * L->speed_limit);
        double spacing = is_ego ? c->ego_spacing : 1 / c->vehicles_density;
        double default_spacing = 12 + 1.0 * speed;
        double offset = spacing * default_spacing * c->spawn_exp;
        double x0;
        if (v > 0) {
            x0 = lane_s(L, s->x[0], s->y[0]);
            for (int j = 1; j < v; j++) x0 = fmax(x0, lane_s(L, s->x[j], s->y[j]));
        } else {
            x0 = 3 * offset;
        }
        x0 += offset * orc_rng_uniform(rng, 0.9, 1.1);
        lane_position(L, x0, 0, &s->x[v], &s->y[v]);
        s->heading[v] = L->heading;
        s->speed[v] = speed;
        s->lane[v] = closest_lane(lanes, c->lanes_count, s->x[v], s->y[v], s->heading[v]);
        s->target_lane[v] = s->lane[v];
        s->target_speed[v] = speed; /* `target_speed or self.speed` */
        s->crashed[v] = 0;
        s->has_impact[v] = 0;
        s->impact_x[v] = s->impact_y[v] = 0;
        s->timer[v] = 0;
        s->delta[v] = 4.0;
        if (is_ego) {
            s->check_collisions[v] = 1;
            if (c->action_type == 0) {
                s->kind[v] = ORC_KIND_MDP;
                s->speed_index[0] = speed_to_index(c, s->target_speed[v]);
                s->target_speed[v] = c->target_speeds[s->speed_index[0]];
            } else {
                s->kind[v] = ORC_KIND_VEHICLE;
                s->speed_index[0] = -1;
            }
        } else {
            s->kind[v] = ORC_KIND_IDM;
            s->check_collisions[v] = c->others_check_collisions;
            s->timer[v] = py_mod((s->x[v] + s->y[v]) * M_PI, c->lane_change_delay);
            s->delta[v] = orc_rng_uniform(rng, c->delta_lo, c->delta_hi);
        }
    }
}

/* ------------------------------------------------------------------ batched driver */

static void bind_env(const OrcHighwayCfg *c, const OrcBatch *b, int e, OrcHighwayState *s) {
    size_t o = (size_t)e * c->n_vehicles;
    s->x = b->x + o;
    s->y = b->y + o;
    s->heading = b->heading + o;
    s->speed = b->speed + o;
    s->target_speed = b->target_speed + o;
    s->timer = b->timer + o;
    s->delta = b->delta + o;
    s->impact_x = b->impact_x + o;
    s->impact_y = b->impact_y + o;
    s->lane = b->lane + o;
    s->target_lane = b->target_lane + o;
    s->kind = b->kind + o;
    s->crashed = b->crashed + o;
    s->has_impact = b->has_impact + o;
    s->check_collisions = b->check_collisions + o;
    s->speed_index = b->speed_index + e;
    s->time = b->time + e;
}

typedef struct {
    const OrcHighwayCfg *c;
    OrcBatch *b;
    const uint8_t *mask;
    const int32_t *action_i;
    const float *action_f;
    float *obs;
    double *reward;
    uint8_t *terminated, *truncated;
    int autoreset, e0, e1, mode;
} Job;

static void *job_run(void *arg) {
    Job *j = (Job *)arg;
    const OrcHighwayCfg *c = j->c;
    size_t obs_sz = (size_t)c->obs_vehicles_count * orc_highway_obs_columns(c);
    for (int e = j->e0; e < j->e1; e++) {
        OrcHighwayState s;
        bind_env(c, j->b, e, &s);
        if (j->mode == 0) { /* reset */
            if (j->mask && !j->mask[e]) continue;
            orc_highway_reset(c, &j->b->rng[e], &s);
            if (j->obs) orc_highway_observe(c, &s, j->obs + obs_sz * e);
        } else {
            double r;
            int32_t te, tr;
            orc_highway_step(c, &s, j->action_i ? j->action_i[e] : 0,
                                 j->action_f ? j->action_f + 2 * (size_t)e : NULL,
                                 j->obs ? j->obs + obs_sz * e : NULL, &r, &te, &tr);
            j->reward[e] = r;
            j->terminated[e] = (uint8_t)te;
            j->truncated[e] = (uint8_t)tr;
            if (j->autoreset && (te || tr)) { /* SameStep: obs becomes the reset obs */
                orc_highway_reset(c, &j->b->rng[e], &s);
                if (j->obs) orc_highway_observe(c, &s, j->obs + obs_sz * e);
            }
        }
    }
    return NULL;
}

static void run_jobs(Job *proto, int n_envs, int threads) {
    if (threads < 1) threads = 1;
    if (threads > n_envs) threads = n_envs > 0 ? n_envs : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    Job *jobs = (Job *)malloc(sizeof(Job) * threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = *proto;
        jobs[t].e0 = (int)((long long)n_envs * t / threads);
        jobs[t].e1 = (int)((long long)n_envs * (t + 1) / threads);
        if (threads == 1)
            job_run(&jobs[t]);
        else
            pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    if (threads > 1)
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

void orc_highway_reset_batch(const OrcHighwayCfg *c, OrcBatch *b, const uint8_t *mask, float *obs,
                             int threads) {
    Job j;
    memset(&j, 0, sizeof(j));
    j.c = c;
    j.b = b;
    j.mask = mask;
    j.obs = obs;
    j.mode = 0;
    run_jobs(&j, b->n_envs, threads);
}

void orc_highway_step_batch(const OrcHighwayCfg *c, OrcBatch *b, const int32_t *action_i,
                                const float *action_f, float *obs, double *reward,
                                uint8_t *terminated, uint8_t *truncated, int autoreset,
                                int threads) {
    Job j;
    memset(&j, 0, sizeof(j));
    j.c = c;
    j.b = b;
    j.action_i = action_i;
    j.action_f = action_f;
    j.obs = obs;
    j.reward = reward;
    j.terminated = terminated;
    j.truncated = truncated;
    j.autoreset = autoreset;
    j.mode = 1;
    run_jobs(&j, b->n_envs, threads);
}
