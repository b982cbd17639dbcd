/*
 * net_oracle.c — scalar CPU restatement of the HighwayEnv hot path on a general road network
 * (roundabout-v0).  TEST INFRASTRUCTURE ONLY — see net_oracle.h.  Same build flags and numpy
 * fused-operation conventions as hwy_oracle.c.  Paths relative to /root/reference/highway_env.
 */
#include "net_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define VEH_LENGTH 5.0
#define VEH_WIDTH 2.0
#define MAX_SPEED 40.0
#define MIN_SPEED (-40.0)
#define LANE_VEHICLE_LENGTH 5.0
static const double TAU_ACC = 0.6, TAU_HEADING = 0.2, TAU_LATERAL = 0.6;
#define TAU_PURSUIT (0.5 * TAU_HEADING)
#define KP_A (1 / TAU_ACC)
#define KP_HEADING (1 / TAU_HEADING)
#define KP_LATERAL (1 / TAU_LATERAL)
#define MAX_STEERING_ANGLE (M_PI / 3)

/* ------------------------------------------------------------------ utils.py */

static inline double dot2(double a0, double a1, double b0, double b1) {
    return fma(a1, b1, a0 * b0); /* np.dot on 2-vectors, see file header */
}
static inline double norm2(double a0, double a1) { return sqrt(dot2(a0, a1, a0, a1)); }
static inline double clipd(double x, double lo, double hi) { /* np.clip */
    return fmin(fmax(x, lo), hi);
}

/* utils.py:50-56 */
static double orc_not_zero(double x) {
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
static double orc_wrap_to_pi(double x) { return py_mod(x + M_PI, 2 * M_PI) - M_PI; }

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
static int orc_rotated_rectangles_intersect(double c1x, double c1y, double l1, double w1, double a1,
                                     double c2x, double c2y, double l2, double w2, double a2) {
    return has_corner_inside(c1x, c1y, l1, w1, a1, c2x, c2y, l2, w2, a2) ||
           has_corner_inside(c2x, c2y, l2, w2, a2, c1x, c1y, l1, w1, a1);
}

/* test entry for the reference's known-answer test (tests/test_utils.py:19-27) */
int net_rotated_rectangles_intersect(double c1x, double c1y, double l1, double w1, double a1, double c2x, double c2y,
                                     double l2, double w2, double a2) {
    return orc_rotated_rectangles_intersect(c1x, c1y, l1, w1, a1, c2x, c2y, l2, w2, a2);
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
static void orc_polygons_intersecting(const double a[5][2], const double b[5][2], double dax, double day,
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


/* ------------------------------------------------------------------ lanes (road/lane.py) */

/* local_coordinates: StraightLane :205-209, SineLane :285-289, CircularLane :351-358 */
void net_lane_local(const NetLane *L, double x, double y, double *s, double *lat) {
    if (L->type == NET_LANE_CIRCULAR) {
        double ddx = x - L->cx, ddy = y - L->cy;
        double phi = atan2(ddy, ddx);
        phi = L->start_phase + orc_wrap_to_pi(phi - L->start_phase);
        double r = norm2(ddx, ddy);
        *s = L->direction * (phi - L->start_phase) * L->radius;
        *lat = L->direction * (L->radius - r);
        return;
    }
    double ddx = x - L->sx, ddy = y - L->sy;
    double lon = dot2(ddx, ddy, L->dx, L->dy);
    double la = dot2(ddx, ddy, L->lx, L->ly);
    if (L->type == NET_LANE_SINE) la = la - L->amplitude * sin(L->pulsation * lon + L->phase);
    *s = lon;
    *lat = la;
}
/* position: StraightLane :192-197, SineLane :268-273, CircularLane :338-342 */
void net_lane_position(const NetLane *L, double s, double lat, double *x, double *y) {
    if (L->type == NET_LANE_CIRCULAR) {
        double phi = L->direction * s / L->radius + L->start_phase;
        double rr = L->radius - lat * L->direction;
        *x = L->cx + rr * cos(phi);
        *y = L->cy + rr * sin(phi);
        return;
    }
    if (L->type == NET_LANE_SINE) lat = lat + L->amplitude * sin(L->pulsation * s + L->phase);
    *x = (L->sx + s * L->dx) + lat * L->lx;
    *y = (L->sy + s * L->dy) + lat * L->ly;
}
/* heading_at: StraightLane :199-200, SineLane :275-280, CircularLane :344-347 */
double net_lane_heading_at(const NetLane *L, double s) {
    if (L->type == NET_LANE_CIRCULAR) {
        double phi = L->direction * s / L->radius + L->start_phase;
        return phi + M_PI / 2 * L->direction;
    }
    if (L->type == NET_LANE_SINE)
        return L->heading + atan(L->amplitude * L->pulsation * cos(L->pulsation * s + L->phase));
    return L->heading;
}
static inline double lane_s(const NetLane *L, double x, double y) {
    double s, lat;
    net_lane_local(L, x, y, &s, &lat);
    return s;
}
/* :80-102 */
static inline int lane_on_lane(const NetLane *L, double s, double lat, double margin) {
    return fabs(lat) <= L->width / 2 + margin && -LANE_VEHICLE_LENGTH <= s &&
           s < L->length + LANE_VEHICLE_LENGTH;
}
/* :104-118 */
static inline int lane_reachable(const NetLane *L, double x, double y) {
    if (L->forbidden) return 0;
    double s, lat;
    net_lane_local(L, x, y, &s, &lat);
    return fabs(lat) <= 2 * L->width && 0 <= s && s < L->length + LANE_VEHICLE_LENGTH;
}
/* :120-125 after_end (longitudinal recomputed) */
static inline int lane_after_end(const NetLane *L, double x, double y) {
    return lane_s(L, x, y) > L->length - LANE_VEHICLE_LENGTH / 2;
}
/* :127-130 distance */
static inline double lane_distance(const NetLane *L, double x, double y) {
    double s, r;
    net_lane_local(L, x, y, &s, &r);
    return fabs(r) + fmax(s - L->length, 0) + fmax(0 - s, 0);
}
/* :132-147 distance_with_heading / local_angle */
static inline double lane_distance_with_heading(const NetLane *L, double x, double y, double h) {
    double s, r;
    net_lane_local(L, x, y, &s, &r);
    double angle = fabs(orc_wrap_to_pi(h - net_lane_heading_at(L, s)));
    return fabs(r) + fmax(s - L->length, 0) + fmax(0 - s, 0) + 1.0 * angle;
}
/* road/road.py:55-71 get_closest_lane_index */
int net_closest_lane(const NetGraph *g, double x, double y, double h) {
    int best = 0;
    double bd = 0;
    for (int l = 0; l < g->n_lanes; l++) {
        double d = lane_distance_with_heading(&g->lanes[l], x, y, h);
        if (l == 0 || d < bd) {
            bd = d;
            best = l;
        }
    }
    return best;
}

/* ------------------------------------------------------------------ world */

typedef struct {
    const NetGraph *g;
    const NetCfg *c;
    NetState *s;
    double *act_steer, *act_accel;
    int V;
} World;

static inline int world_count(const NetCfg *c, const NetState *s) { return s->count ? s->count[0] : c->n_vehicles; }

#define RT_FROM(e) ((e)&0xff)
#define RT_TO(e) (((e) >> 8) & 0xff)
#define RT_ID(e) ((((e) >> 16) & 0xff) - 1) /* -1 = None */

static inline const NetLane *LANE(const World *w, int idx) { return &w->g->lanes[idx]; }

/* lane table index of road (from, to): -1 if the road does not exist */
static int road_first(const NetGraph *g, int from, int to) {
    for (int k = 0; k < g->succ_count[from]; k++) {
        int f = g->succ[from][k];
        if (g->lanes[f].to_node == to) return f;
    }
    return -1;
}

/* vehicle/objects.py:183-198 */
static inline double lane_distance_to(const World *w, int self, int other) {
    const NetLane *L = LANE(w, w->s->lane[self]);
    return lane_s(L, w->s->x[other], w->s->y[other]) - lane_s(L, w->s->x[self], w->s->y[self]);
}

/* road/road.py:483-547 neighbour_vehicles.  With neighbour_vehicles_connected_lanes (:509-529) the lanes
 * that continue `lane_index` (every road leaving its end node: lane _id, or lane 0 when that road has fewer
 * lanes) are searched with offset +length, and the lanes leading into its start node (from-nodes in graph
 * insertion order) with offset -their length; a vehicle is counted on the FIRST lane of that list it is on. */
static void neighbour_vehicles(const World *w, int veh, int lane_idx, int *front, int *rear) {
    const NetLane *L = LANE(w, lane_idx);
    double s = lane_s(L, w->s->x[veh], w->s->y[veh]);
    double s_front = 0, s_rear = 0;
    int v_front = -1, v_rear = -1;
    int lanes[1 + NET_MAX_SUCC + NET_MAX_NODES];
    double offsets[1 + NET_MAX_SUCC + NET_MAX_NODES];
    int n_l = 0;
    lanes[n_l] = lane_idx;
    offsets[n_l++] = 0.0;
    if (w->c->connected_lanes) {
        const NetGraph *g = w->g;
        for (int k = 0; k < g->succ_count[L->to_node]; k++) {
            const NetLane *N0 = &g->lanes[g->succ[L->to_node][k]];
            lanes[n_l] = g->succ[L->to_node][k] + (L->lane_id < N0->road_count ? L->lane_id : 0);
            offsets[n_l++] = L->length;
        }
        for (int l = 0; l < g->n_lanes; l++) { /* table order == graph.values() order of the from-nodes */
            const NetLane *P0 = &g->lanes[l];
            if (P0->lane_id != 0 || P0->to_node != L->from_node) continue;
            int pl = l + (L->lane_id < P0->road_count ? L->lane_id : 0);
            lanes[n_l] = pl;
            offsets[n_l++] = -g->lanes[pl].length;
        }
    }
    for (int v = 0; v < w->V; v++) {
        if (v == veh) continue;
        for (int k = 0; k < n_l; k++) {
            const NetLane *SL = LANE(w, lanes[k]);
            double s_v, lat_v;
            net_lane_local(SL, w->s->x[v], w->s->y[v], &s_v, &lat_v);
            if (!lane_on_lane(SL, s_v, lat_v, 1.0)) continue;
            s_v += offsets[k];
            if (s <= s_v && (v_front < 0 || s_v <= s_front)) {
                s_front = s_v;
                v_front = v;
            }
            if (s_v < s && (v_rear < 0 || s_v > s_rear)) {
                s_rear = s_v;
                v_rear = v;
            }
            break; /* matched on this lane */
        }
    }
    *front = v_front;
    *rear = v_rear;
}

/* test entry: Road.neighbour_vehicles(vehicle, lane_index) for the reference's own known-answer tests
 * (tests/road/test_neighbour_vehicles.py) */
void net_neighbours(const NetGraph *g, const NetCfg *c, const NetState *s, int veh, int lane_idx, int32_t *front,
                    int32_t *rear) {
    World w;
    w.g = g;
    w.c = c;
    w.s = (NetState *)s;
    w.V = world_count(c, s);
    int f, r;
    neighbour_vehicles(&w, veh, lane_idx, &f, &r);
    *front = f;
    *rear = r;
}

/* vehicle/behavior.py:192-217 */
static double desired_gap(const World *w, int ego, int front) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    double ab = -c->comfort_acc_max * c->comfort_acc_min;
    double ce = cos(s->heading[ego]), se = sin(s->heading[ego]);
    double cf = cos(s->heading[front]), sf = sin(s->heading[front]);
    double dvx = s->speed[ego] * ce - s->speed[front] * cf;
    double dvy = s->speed[ego] * se - s->speed[front] * sf;
    double dv = dot2(dvx, dvy, ce, se);
    return c->distance_wanted + s->speed[ego] * c->time_wanted + s->speed[ego] * dv / (2 * sqrt(ab));
}

/* vehicle/behavior.py:150-190 */
static double idm_acceleration(const World *w, int self_, int ego, int front) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    if (ego < 0 || s->kind[ego] == NET_KIND_OBSTACLE) return 0; /* `not isinstance(ego_vehicle, Vehicle)` (behavior.py:171) */
    /* getattr(ego_vehicle, "target_speed", 0): a plain Vehicle has none (behavior.py:172) */
    double ego_ts = s->kind[ego] == NET_KIND_VEHICLE ? 0.0 : s->target_speed[ego];
    double ego_target_speed = clipd(ego_ts, 0, LANE(w, s->lane[ego])->speed_limit);
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

/* vehicle/controller.py:145-187 */
static double steering_control(const World *w, int v, int target_lane) {
    const NetState *s = w->s;
    const NetLane *L = LANE(w, target_lane);
    double lc_s, lc_lat;
    net_lane_local(L, s->x[v], s->y[v], &lc_s, &lc_lat);
    double tau = (s->kind[v] == NET_KIND_MDP && w->c->ego_pursuit_tau > 0) ? w->c->ego_pursuit_tau : TAU_PURSUIT;
    double lane_next_coords = lc_s + s->speed[v] * tau;
    double lane_future_heading = net_lane_heading_at(L, lane_next_coords);
    double lateral_speed_command = -KP_LATERAL * lc_lat;
    double heading_command = asin(clipd(lateral_speed_command / orc_not_zero(s->speed[v]), -1, 1));
    double heading_ref = lane_future_heading + clipd(heading_command, -M_PI / 4, M_PI / 4);
    double heading_rate_command = KP_HEADING * orc_wrap_to_pi(heading_ref - s->heading[v]);
    double slip_angle =
        asin(clipd(VEH_LENGTH / 2 / orc_not_zero(s->speed[v]) * heading_rate_command, -1, 1));
    double steering_angle = atan(2 * tan(slip_angle));
    return clipd(steering_angle, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
}

/* road/road.py:138-157 next_lane_given_next_road; next_id < 0 == None */
static int next_lane_given_next_road(const NetGraph *g, int cur, int next_first, int next_id,
                                     double px, double py, double *dist) {
    const NetLane *C = &g->lanes[cur];
    int n_next = g->lanes[next_first].road_count;
    if (C->road_count == n_next) {
        if (next_id < 0) next_id = C->lane_id;
    } else {
        int best = 0;
        double bd = 0;
        for (int l = 0; l < n_next; l++) {
            double d = lane_distance(&g->lanes[next_first + l], px, py);
            if (l == 0 || d < bd) {
                bd = d;
                best = l;
            }
        }
        next_id = best;
    }
    *dist = lane_distance(&g->lanes[next_first + next_id], px, py);
    return next_id;
}

/* road/road.py:73-136 next_lane; mutates the vehicle's route (pop(0)) */
static int next_lane(World *w, int v, int cur) {
    const NetGraph *g = w->g;
    NetState *s = w->s;
    const NetLane *C = &g->lanes[cur];
    int32_t *route = s->route + (size_t)v * NET_MAX_ROUTE;
    int *rlen = &s->route_len[v];
    int next_first = -1, next_id = -1;
    if (*rlen > 0) {
        if (RT_FROM(route[0]) == C->from_node && RT_TO(route[0]) == C->to_node) {
            for (int k = 1; k < *rlen; k++) route[k - 1] = route[k];
            (*rlen)--;
        }
        if (*rlen > 0 && RT_FROM(route[0]) == C->to_node) {
            next_first = road_first(g, RT_FROM(route[0]), RT_TO(route[0]));
            next_id = RT_ID(route[0]);
        }
        /* else: logger.warning only */
    }
    double lon, lat, px, py;
    net_lane_local(C, s->x[v], s->y[v], &lon, &lat);
    net_lane_position(C, lon, 0, &px, &py);
    if (next_first < 0) {
        int n_succ = g->succ_count[C->to_node];
        if (n_succ == 0) return cur; /* KeyError: graph[_to] */
        int best_first = -1, best_id = -1;
        double bd = 0;
        for (int k = 0; k < n_succ; k++) {
            int nf = g->succ[C->to_node][k];
            double d;
            int nid = next_lane_given_next_road(g, cur, nf, next_id, px, py, &d);
            if (k == 0 || d < bd) { /* min(): first minimum */
                bd = d;
                best_first = nf;
                best_id = nid;
            }
        }
        return best_first + best_id;
    }
    double d;
    next_id = next_lane_given_next_road(g, cur, next_first, next_id, px, py, &d);
    return next_first + next_id;
}

/* vehicle/controller.py:135-143 */
static void follow_road(World *w, int v) {
    NetState *s = w->s;
    if (lane_after_end(LANE(w, s->target_lane[v]), s->x[v], s->y[v]))
        s->target_lane[v] = next_lane(w, v, s->target_lane[v]);
}

static inline int isign(int a) { return (a > 0) - (a < 0); }

/* vehicle/behavior.py:265-324 */
static int mobil(const World *w, int v, int lane_index) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    int new_preceding, new_following;
    neighbour_vehicles(w, v, lane_index, &new_preceding, &new_following);
    double new_following_a = idm_acceleration(w, v, new_following, new_preceding);
    double new_following_pred_a = idm_acceleration(w, v, new_following, v);
    if (new_following_pred_a < -c->lane_change_max_braking_imposed) return 0;
    int old_preceding, old_following;
    neighbour_vehicles(w, v, s->lane[v], &old_preceding, &old_following);
    double self_pred_a = idm_acceleration(w, v, v, new_preceding);
    const int32_t *route = s->route + (size_t)v * NET_MAX_ROUTE;
    if (s->route_len[v] > 0 && RT_ID(route[0]) >= 0) {
        int tid = w->g->lanes[s->target_lane[v]].lane_id;
        int cid = w->g->lanes[lane_index].lane_id;
        if (isign(cid - tid) != isign(RT_ID(route[0]) - tid)) return 0;
        if (self_pred_a < -c->lane_change_max_braking_imposed) return 0;
    } else {
        double self_a = idm_acceleration(w, v, v, old_preceding);
        double old_following_a = idm_acceleration(w, v, old_following, v);
        double old_following_pred_a = idm_acceleration(w, v, old_following, old_preceding);
        double jerk = self_pred_a - self_a +
                      c->politeness * (new_following_pred_a - new_following_a +
                                       old_following_pred_a - old_following_a);
        if (jerk < c->lane_change_min_acc_gain) return 0;
    }
    return 1;
}

/* vehicle/behavior.py:219-263 */
static void change_lane_policy(World *w, int v) {
    const NetCfg *c = w->c;
    NetState *s = w->s;
    const NetGraph *g = w->g;
    if (s->lane[v] != s->target_lane[v]) {
        const NetLane *A = &g->lanes[s->lane[v]], *B = &g->lanes[s->target_lane[v]];
        if (A->from_node == B->from_node && A->to_node == B->to_node) {
            for (int o = 0; o < w->V; o++) {
                if (o != v && s->lane[o] != s->target_lane[v] && s->target_lane[o] == s->target_lane[v]) {
                    double d = lane_distance_to(w, v, o);
                    double d_star = desired_gap(w, v, o);
                    if (0 < d && d < d_star) {
                        s->target_lane[v] = s->lane[v];
                        break;
                    }
                }
            }
        }
        return;
    }
    if (!(c->lane_change_delay < s->timer[v])) return;
    s->timer[v] = 0;
    const NetLane *A = &g->lanes[s->lane[v]];
    int cand[2], nc = 0; /* road/road.py:200-211 side_lanes */
    if (A->lane_id > 0) cand[nc++] = s->lane[v] - 1;
    if (A->lane_id < A->road_count - 1) cand[nc++] = s->lane[v] + 1;
    for (int k = 0; k < nc; k++) {
        if (!lane_reachable(&g->lanes[cand[k]], s->x[v], s->y[v])) continue;
        if (fabs(s->speed[v]) < 1) continue;
        if (mobil(w, v, cand[k])) s->target_lane[v] = cand[k];
    }
}

/* vehicle/behavior.py:93-137 */
static void idm_act(World *w, int v) {
    const NetCfg *c = w->c;
    NetState *s = w->s;
    if (s->crashed[v]) return;
    follow_road(w, v);
    if (!(s->no_lane_change && s->no_lane_change[v])) change_lane_policy(w, v); /* behavior.py:104-105 */
    double steering = steering_control(w, v, s->target_lane[v]);
    steering = clipd(steering, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
    int front, rear;
    neighbour_vehicles(w, v, s->lane[v], &front, &rear);
    double acc = idm_acceleration(w, v, v, front);
    if (s->lane[v] != s->target_lane[v]) {
        neighbour_vehicles(w, v, s->target_lane[v], &front, &rear);
        acc = fmin(acc, idm_acceleration(w, v, v, front));
    }
    w->act_steer[v] = steering;
    w->act_accel[v] = clipd(acc, -c->acc_max, c->acc_max);
}

/* vehicle/controller.py:89-133; label 0 LANE_LEFT, 2 LANE_RIGHT, else none */
static void controlled_act(World *w, int v, int label) {
    NetState *s = w->s;
    const NetGraph *g = w->g;
    follow_road(w, v);
    if (label == 0 || label == 2) {
        const NetLane *T = &g->lanes[s->target_lane[v]];
        int id = T->lane_id + (label == 2 ? 1 : -1);
        if (id < 0) id = 0;
        if (id > T->road_count - 1) id = T->road_count - 1;
        int cand = T->road_first + id;
        if (lane_reachable(&g->lanes[cand], s->x[v], s->y[v])) s->target_lane[v] = cand;
    }
    double steering = steering_control(w, v, s->target_lane[v]);
    w->act_steer[v] = clipd(steering, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
    w->act_accel[v] = KP_A * (s->target_speed[v] - s->speed[v]);
}

/* vehicle/controller.py:326-344 */
static int speed_to_index(const NetCfg *c, double speed) {
    int n = c->n_target_speeds;
    double x = (speed - c->target_speeds[0]) / (c->target_speeds[n - 1] - c->target_speeds[0]);
    return (int)clipd(rint(x * (n - 1)), 0, n - 1);
}

/* vehicle/controller.py:295-315; action.py:204 labels */
static void mdp_act_agent(World *w, int v, int action, int agent) {
    const NetCfg *c = w->c;
    NetState *s = w->s;
    if (action == 3 || action == 4) {
        int idx = speed_to_index(c, s->speed[v]) + (action == 3 ? 1 : -1);
        if (idx < 0) idx = 0;
        if (idx > c->n_target_speeds - 1) idx = c->n_target_speeds - 1;
        s->speed_index[agent] = idx;
        s->target_speed[v] = c->target_speeds[idx];
        controlled_act(w, v, -1);
    } else {
        controlled_act(w, v, action);
    }
}

static void mdp_act(World *w, int v, int action) { mdp_act_agent(w, v, action, 0); }

/* the k-th controlled vehicle (env.controlled_vehicles order == list order of the MDP vehicles); -1: none */
static int agent_vehicle(const World *w, int k) {
    for (int v = 0; v < w->V; v++)
        if ((w->s->kind[v] == NET_KIND_MDP || w->s->kind[v] == NET_KIND_VEHICLE) && k-- == 0) return v;
    return -1;
}

/* index of the controlled vehicle: first MDP vehicle of the list */
static int ego_index(const World *w) {
    for (int v = 0; v < w->V; v++)
        if (w->s->kind[v] == NET_KIND_MDP || w->s->kind[v] == NET_KIND_VEHICLE) return v;
    return 0;
}

static void road_act(World *w) {
    for (int v = 0; v < w->V; v++) {
        if (w->s->kind[v] == NET_KIND_OBSTACLE) continue; /* road.objects do not act (road.py:464-467) */
        if (w->s->kind[v] == NET_KIND_VEHICLE) continue;  /* Vehicle.act(None): keeps its action (kinematics.py:119-128) */
        if (w->s->kind[v] == NET_KIND_IDM)
            idm_act(w, v);
        else
            controlled_act(w, v, -1);
    }
}

/* BicycleVehicle.derivative_func (vehicle/dynamics.py:73-111) on state (x, y, heading, speed, lateral_speed,
 * yaw_rate) with the action's steering / acceleration */
#define BIKE_MASS 1.0
#define BIKE_LENGTH_A (VEH_LENGTH / 2)
#define BIKE_LENGTH_B (VEH_LENGTH / 2)
#define BIKE_INERTIA_Z (1.0 / 12 * BIKE_MASS * (VEH_LENGTH * VEH_LENGTH + VEH_WIDTH * VEH_WIDTH))
#define BIKE_FRICTION_FRONT (15.0 * BIKE_MASS)
#define BIKE_FRICTION_REAR (15.0 * BIKE_MASS)
static void bicycle_derivative(const double st[6], double steering, double acceleration, double d[6]) {
    const double heading = st[2], speed = st[3], lateral_speed = st[4], yaw_rate = st[5];
    const double delta_f = steering, delta_r = 0;
    double theta_vf = atan2(lateral_speed + BIKE_LENGTH_A * yaw_rate, speed);
    double theta_vr = atan2(lateral_speed - BIKE_LENGTH_B * yaw_rate, speed);
    double f_yf = 2 * BIKE_FRICTION_FRONT * (delta_f - theta_vf);
    double f_yr = 2 * BIKE_FRICTION_REAR * (delta_r - theta_vr);
    if (fabs(speed) < 1) { /* low speed dynamics: damping of lateral speed and yaw rate */
        f_yf = -BIKE_MASS * lateral_speed - BIKE_INERTIA_Z / BIKE_LENGTH_A * yaw_rate;
        f_yr = -BIKE_MASS * lateral_speed + BIKE_INERTIA_Z / BIKE_LENGTH_A * yaw_rate;
    }
    double d_lateral_speed = 1 / BIKE_MASS * (f_yf + f_yr) - yaw_rate * speed;
    double d_yaw_rate = 1 / BIKE_INERTIA_Z * (BIKE_LENGTH_A * f_yf - BIKE_LENGTH_B * f_yr);
    double c = cos(heading), sn = sin(heading);
    d[0] = c * speed + (-sn) * lateral_speed; /* R @ [speed, lateral_speed] */
    d[1] = sn * speed + c * lateral_speed;
    d[2] = yaw_rate;
    d[3] = acceleration;
    d[4] = d_lateral_speed;
    d[5] = d_yaw_rate;
}
/* dynamics.py:13-30 rk4 and BicycleVehicle.step (:142-150) on an explicit state; clip_actions (:153-161) first */
static void bicycle_advance(double st[6], int crashed, double *steering, double *acceleration, double dt) {
    if (crashed) { /* Vehicle.clip_actions (kinematics.py:155-168) */
        *steering = 0;
        *acceleration = -1.0 * st[3];
    }
    if (st[3] > MAX_SPEED)
        *acceleration = fmin(*acceleration, 1.0 * (MAX_SPEED - st[3]));
    else if (st[3] < MIN_SPEED)
        *acceleration = fmax(*acceleration, 1.0 * (MIN_SPEED - st[3]));
    *steering = clipd(*steering, -M_PI / 2, M_PI / 2);
    st[5] = clipd(st[5], -2 * M_PI, 2 * M_PI); /* MAX_ANGULAR_SPEED */
    double f1[6], f2[6], f3[6], f4[6], tmp[6];
    bicycle_derivative(st, *steering, *acceleration, f1);
    for (int k = 0; k < 6; k++) tmp[k] = st[k] + (f1[k] * (dt / 2));
    bicycle_derivative(tmp, *steering, *acceleration, f2);
    for (int k = 0; k < 6; k++) tmp[k] = st[k] + (f2[k] * (dt / 2));
    bicycle_derivative(tmp, *steering, *acceleration, f3);
    for (int k = 0; k < 6; k++) tmp[k] = st[k] + (f3[k] * dt);
    bicycle_derivative(tmp, *steering, *acceleration, f4);
    for (int k = 0; k < 6; k++) st[k] = st[k] + (dt / 6) * (f1[k] + (2 * f2[k]) + (2 * f3[k]) + f4[k]);
}
/* Vehicle.step on an explicit state (kinematics.py:130-153), the impact handled by the caller */
static void kinematic_advance(double st[4], int crashed, double *steering, double *acceleration, double dt) {
    if (crashed) {
        *steering = 0;
        *acceleration = -1.0 * st[3];
    }
    if (st[3] > MAX_SPEED)
        *acceleration = fmin(*acceleration, 1.0 * (MAX_SPEED - st[3]));
    else if (st[3] < MIN_SPEED)
        *acceleration = fmax(*acceleration, 1.0 * (MIN_SPEED - st[3]));
    double beta = atan(1.0 / 2 * tan(*steering));
    double vx = st[3] * cos(st[2] + beta), vy = st[3] * sin(st[2] + beta);
    st[0] += vx * dt;
    st[1] += vy * dt;
    st[2] += st[3] * sin(beta) / (VEH_LENGTH / 2) * dt;
    st[3] += *acceleration * dt;
}

/* vehicle/kinematics.py:130-177 (+ behavior.py:139-148) */
static void vehicle_step(World *w, int v, double dt) {
    NetState *s = w->s;
    if (s->kind[v] == NET_KIND_OBSTACLE) return; /* only road.vehicles step (road.py:475-476) */
    if (s->kind[v] == NET_KIND_VEHICLE && w->c->dynamical) { /* BicycleVehicle.step: no impact handling of its own */
        double st[6] = {s->x[v], s->y[v], s->heading[v], s->speed[v], s->lat_speed[v], s->yaw_rate[v]};
        /* clip_actions clips yaw_rate in place before the state is read */
        bicycle_advance(st, s->crashed[v], &w->act_steer[v], &w->act_accel[v], dt);
        s->x[v] = st[0];
        s->y[v] = st[1];
        s->heading[v] = st[2];
        s->speed[v] = st[3];
        s->lat_speed[v] = st[4];
        s->yaw_rate[v] = st[5];
        s->lane[v] = s->target_lane[v] = net_closest_lane(w->g, s->x[v], s->y[v], s->heading[v]);
        return;
    }
    if (s->kind[v] == NET_KIND_IDM) s->timer[v] += dt;
    if (s->crashed[v]) {
        w->act_steer[v] = 0;
        w->act_accel[v] = -1.0 * s->speed[v];
    }
    if (s->speed[v] > MAX_SPEED)
        w->act_accel[v] = fmin(w->act_accel[v], 1.0 * (MAX_SPEED - s->speed[v]));
    else if (s->speed[v] < MIN_SPEED)
        w->act_accel[v] = fmax(w->act_accel[v], 1.0 * (MIN_SPEED - s->speed[v]));
    double beta = atan(1.0 / 2 * tan(w->act_steer[v]));
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
    s->lane[v] = net_closest_lane(w->g, s->x[v], s->y[v], s->heading[v]);
    if (s->kind[v] == NET_KIND_VEHICLE) s->target_lane[v] = s->lane[v]; /* schema: no target lane of its own */
}

static double object_length(const NetState *s, int v) { /* RoadObject.LENGTH 2 (objects.py:25), Vehicle 5 */
    return s->kind[v] == NET_KIND_OBSTACLE ? 2.0 : VEH_LENGTH;
}
static void polygon(const NetState *s, int v, double p[5][2]) {
    const double len = object_length(s, v);
    const double loc[4][2] = {{-len / 2, -VEH_WIDTH / 2},
                              {-len / 2, +VEH_WIDTH / 2},
                              {+len / 2, +VEH_WIDTH / 2},
                              {+len / 2, -VEH_WIDTH / 2}};
    double c = cos(s->heading[v]), sn = sin(s->heading[v]);
    for (int k = 0; k < 4; k++) {
        p[k][0] = (c * loc[k][0] + (-sn) * loc[k][1]) + s->x[v];
        p[k][1] = (sn * loc[k][0] + c * loc[k][1]) + s->y[v];
    }
    p[4][0] = p[0][0];
    p[4][1] = p[0][1];
}

/* vehicle/objects.py:92-138 */
static void handle_collisions(World *w, int a, int b, double dt) {
    NetState *s = w->s;
    if (!(s->check_collisions[a] || s->check_collisions[b])) return;
    if (s->kind[a] == NET_KIND_OBSTACLE) return; /* Road.step: only vehicles call handle_collisions (road.py:477-481) */
    double la = object_length(s, a), lb = object_length(s, b);
    double diag_a = sqrt(la * la + VEH_WIDTH * VEH_WIDTH), diag_b = sqrt(lb * lb + VEH_WIDTH * VEH_WIDTH);
    double dist = norm2(s->x[b] - s->x[a], s->y[b] - s->y[a]);
    if (dist > (diag_a + diag_b) / 2 + s->speed[a] * dt) return;
    double pa[5][2], pb[5][2], tr[2];
    polygon(s, a, pa);
    polygon(s, b, pb);
    double ca = cos(s->heading[a]), sa = sin(s->heading[a]);
    double cb = cos(s->heading[b]), sb = sin(s->heading[b]);
    int inter, will;
    orc_polygons_intersecting(pa, pb, s->speed[a] * ca * dt, s->speed[a] * sa * dt,
                              s->speed[b] * cb * dt, s->speed[b] * sb * dt, &inter, &will, tr);
    if (will) {
        if (s->kind[b] == NET_KIND_OBSTACLE) { /* objects.py:106-107: the whole transition goes to the vehicle */
            s->impact_x[a] = tr[0];
            s->impact_y[a] = tr[1];
            s->has_impact[a] = 1;
        } else {
            s->impact_x[a] = tr[0] / 2;
            s->impact_y[a] = tr[1] / 2;
            s->has_impact[a] = 1;
            s->impact_x[b] = -tr[0] / 2;
            s->impact_y[b] = -tr[1] / 2;
            s->has_impact[b] = 1;
        }
    }
    if (inter) {
        s->crashed[a] = 1;
        s->crashed[b] = 1;
    }
}

static void road_step(World *w, double dt) {
    for (int v = 0; v < w->V; v++) vehicle_step(w, v, dt);
    for (int i = 0; i < w->V; i++)
        for (int j = i + 1; j < w->V; j++) handle_collisions(w, i, j, dt);
}

/* ------------------------------------------------------------------ observations */

/* road/road.py:231-276 is_connected_road(l1, l2, route, same_lane=False, depth) with l1 given
 * as (from, to) — lane ids are irrelevant when same_lane is False. */
static int is_connected_road(const NetGraph *g, int f1, int t1, int f2, int t2, const int32_t *route,
                             int rlen, int depth) {
    if ((f2 == f1 && t2 == t1) || (t2 == f1)) return 1; /* is_same_road or is_leading_to_road */
    if (depth > 0) {
        if (rlen > 0 && RT_FROM(route[0]) == f1 && RT_TO(route[0]) == t1)
            return is_connected_road(g, f1, t1, f2, t2, route + 1, rlen - 1, depth);
        if (rlen > 0 && RT_FROM(route[0]) == t1)
            return is_connected_road(g, RT_FROM(route[0]), RT_TO(route[0]), f2, t2, route + 1,
                                     rlen - 1, depth - 1);
        int any = 0;
        for (int k = 0; k < g->succ_count[t1]; k++) {
            int nf = g->succ[t1][k];
            if (is_connected_road(g, t1, g->lanes[nf].to_node, f2, t2, route, rlen, depth - 1)) any = 1;
        }
        return any;
    }
    return 0;
}

/* envs/common/finite_mdp.py:104-163 compute_ttc_grid + observation.py:128-152 */
static void observe_ttc_from(const World *w, int ego, int si_e, float *obs) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const NetGraph *g = w->g;
    const NetLane *EL = &g->lanes[s->lane[ego]];
    int n_speeds = c->n_target_speeds, n_lanes = EL->road_count;
    int n_t = (int)(c->ttc_horizon / (1.0 / c->policy_frequency));
    double tq = 1.0 / c->policy_frequency;
    double *grid = (double *)calloc((size_t)n_speeds * n_lanes * n_t, sizeof(double));
    double ce = cos(s->heading[ego]), se = sin(s->heading[ego]);
    const int32_t *route = s->route + (size_t)ego * NET_MAX_ROUTE;
    for (int si = 0; si < n_speeds; si++) {
        double ego_speed = c->target_speeds[si];
        for (int o = 0; o < w->V; o++) {
            if (o == ego || ego_speed == s->speed[o] || s->kind[o] == NET_KIND_OBSTACLE) continue;
            double margin = VEH_LENGTH / 2 + VEH_LENGTH / 2;
            const double ms[3] = {0, -margin, margin}, costs[3] = {1, 0.5, 0.5};
            const NetLane *OL = &g->lanes[s->lane[o]];
            for (int k = 0; k < 3; k++) {
                double distance = lane_distance_to(w, ego, o) + ms[k];
                double other_projected_speed =
                    s->speed[o] * dot2(cos(s->heading[o]), sin(s->heading[o]), ce, se);
                double ttc = distance / orc_not_zero(ego_speed - other_projected_speed);
                if (ttc < 0) continue;
                if (!is_connected_road(g, EL->from_node, EL->to_node, OL->from_node, OL->to_node, route,
                                       s->route_len[ego], 3))
                    continue;
                int l0 = 0, l1 = n_lanes; /* all lanes */
                if (OL->road_count == EL->road_count) {
                    l0 = OL->lane_id;
                    l1 = l0 + 1;
                }
                int times[2] = {(int)(ttc / tq), (int)ceil(ttc / tq)};
                for (int q = 0; q < 2; q++) {
                    int t = times[q];
                    if (0 <= t && t < n_t)
                        for (int l = l0; l < l1; l++) {
                            double *cell = &grid[((size_t)si * n_lanes + l) * n_t + t];
                            *cell = fmax(*cell, costs[k]);
                        }
                }
            }
        }
    }
    /* pad lanes with ones, crop 3 around the ego lane; repeat first/last speed rows, crop 3 */
    int ego_lane_id = EL->lane_id;
    for (int a = 0; a < 3; a++) {
        int vrow = n_speeds + si_e - 1 + a; /* index into the repeated array */
        /* repeated rows: row0 x (1+n_speeds), middle rows x1, last x (1+n_speeds) */
        int src;
        if (vrow < 1 + n_speeds)
            src = 0;
        else if (vrow < 1 + n_speeds + (n_speeds - 2))
            src = 1 + (vrow - (1 + n_speeds));
        else
            src = n_speeds - 1;
        if (n_speeds == 1) src = 0;
        for (int b = 0; b < 3; b++) {
            int lcol = n_lanes + ego_lane_id - 1 + b; /* index into [ones | grid | ones] */
            for (int t = 0; t < n_t; t++) {
                double val;
                if (lcol < n_lanes || lcol >= 2 * n_lanes)
                    val = 1.0;
                else
                    val = grid[((size_t)src * n_lanes + (lcol - n_lanes)) * n_t + t];
                obs[(a * 3 + b) * n_t + t] = (float)val;
            }
        }
    }
    free(grid);
}

static void observe_ttc(const World *w, float *obs) { observe_ttc_from(w, ego_index(w), w->s->speed_index[0], obs); }
void net_observe_ttc_from(const NetGraph *g, const NetCfg *c, const NetState *s, int V, int ego, int speed_index,
                          float *obs) {
    World w;
    w.g = g;
    w.c = c;
    w.s = (NetState *)s;
    w.V = V;
    observe_ttc_from(&w, ego, speed_index, obs);
}

/* envs/common/observation.py:234-276 with explicit features_range / absolute; optional
 * cos_h, sin_h columns (vehicle/kinematics.py:247-248), which have no features_range entry */
static double vehicle_feature(const NetGraph *g, const NetState *s, int v, int origin, int feat, int observe_intentions);
static void observe_kinematics_from(const World *w, int ego, float *obs);
static void observe_kinematics(const World *w, float *obs) { observe_kinematics_from(w, ego_index(w), obs); }
/* KinematicObservation.observe (observation.py:234-276) with a configured column list (any Vehicle.to_dict key,
 * vehicle/kinematics.py:237-261) and per-column ranges: cfg->obs_n_feat > 0 */
static void observe_kinematics_features(const World *w, int ego, float *obs) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const int K = c->obs_vehicles_count, V = w->V, F = c->obs_n_feat;
    int *cand = (int *)malloc(sizeof(int) * (V > 0 ? V : 1));
    double *key = (double *)malloc(sizeof(double) * (V > 0 ? V : 1));
    int nc = 0;
    for (int v = 0; v < V; v++) { /* close_objects_to, as observe_kinematics_from */
        if (!(norm2(s->x[v] - s->x[ego], s->y[v] - s->y[ego]) < c->perception_distance)) continue;
        if (v == ego) continue;
        double d = lane_distance_to(w, ego, v);
        if (!((c->obs_see_behind && s->kind[v] != NET_KIND_OBSTACLE) || -2 * VEH_LENGTH < d)) continue;
        cand[nc] = v;
        key[nc] = fabs(d);
        nc++;
    }
    for (int i = 1; i < nc; i++) { /* stable insertion sort == sorted(key=...) */
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
    for (int k = 0; k < K * F; k++) obs[k] = 0.0f;
    int n_rows = 1 + (nc < K - 1 ? nc : K - 1);
    for (int row = 0; row < n_rows; row++) {
        const int v = row == 0 ? ego : cand[row - 1];
        const int origin = (row == 0 || c->obs_absolute) ? -1 : ego;
        for (int col = 0; col < F; col++) {
            double val = vehicle_feature(w->g, s, v, origin, c->obs_feat[col], 0);
            if (isnan(val)) val = 0; /* road objects have no such column: NaN in the frame */
            if (c->obs_normalize && c->obs_feat_ranged[col]) {
                val = lmap(val, c->obs_feat_lo[col], c->obs_feat_hi[col], -1, 1);
                if (c->obs_clip) val = clipd(val, -1, 1);
            }
            obs[row * F + col] = (float)val;
        }
    }
    free(cand);
    free(key);
}
static void observe_kinematics_from(const World *w, int ego, float *obs) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    if (c->obs_n_feat > 0) {
        observe_kinematics_features(w, ego, obs);
        return;
    }
    int K = c->obs_vehicles_count, V = w->V;
    const int F = c->obs_features == 7 ? 7 : 5;
    double *rows = (double *)calloc((size_t)K * F, sizeof(double));
    double evx = s->speed[ego] * cos(s->heading[ego]), evy = s->speed[ego] * sin(s->heading[ego]);
    rows[0] = 1;
    rows[1] = s->x[ego];
    rows[2] = s->y[ego];
    rows[3] = evx;
    rows[4] = evy;
    if (c->obs_exit_lane > 0) /* ExitObservation.observe :632-636: ego_dict["x"] = exit_lane.local_coordinates(position)[0] */
        rows[1] = lane_s(&w->g->lanes[c->obs_exit_lane], s->x[ego], s->y[ego]);
    if (F == 7) {
        rows[5] = cos(s->heading[ego]);
        rows[6] = sin(s->heading[ego]);
    }
    int *cand = (int *)malloc(sizeof(int) * (V > 0 ? V : 1));
    double *key = (double *)malloc(sizeof(double) * (V > 0 ? V : 1));
    int nc = 0;
    for (int v = 0; v < V; v++) {
        if (!(norm2(s->x[v] - s->x[ego], s->y[v] - s->y[ego]) < c->perception_distance)) continue;
        if (v == ego) continue;
        double d = lane_distance_to(w, ego, v);
        /* road.py:421-450 close_objects_to: obstacles are always filtered like see_behind=False */
        if (!((c->obs_see_behind && s->kind[v] != NET_KIND_OBSTACLE) || -2 * VEH_LENGTH < d)) continue;
        cand[nc] = v;
        key[nc] = fabs(d);
        nc++;
    }
    for (int i = 1; i < nc; i++) {
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
    int n_rows = 1;
    for (int k = 0; k < nc && k < K - 1; k++) {
        int v = cand[k];
        double *r = rows + F * n_rows;
        r[0] = 1;
        r[1] = s->x[v];
        r[2] = s->y[v];
        r[3] = s->speed[v] * cos(s->heading[v]);
        r[4] = s->speed[v] * sin(s->heading[v]);
        if (!c->obs_absolute) {
            r[1] -= s->x[ego];
            r[2] -= s->y[ego];
            r[3] -= evx;
            r[4] -= evy;
        }
        if (F == 7) {
            r[5] = cos(s->heading[v]);
            r[6] = sin(s->heading[v]);
        }
        n_rows++;
    }
    if (c->obs_normalize) {
        for (int k = 0; k < n_rows; k++) {
            double *r = rows + F * k;
            r[1] = lmap(r[1], c->obs_x_lo, c->obs_x_hi, -1, 1);
            r[2] = lmap(r[2], c->obs_y_lo, c->obs_y_hi, -1, 1);
            r[3] = lmap(r[3], c->obs_vx_lo, c->obs_vx_hi, -1, 1);
            r[4] = lmap(r[4], c->obs_vy_lo, c->obs_vy_hi, -1, 1);
            if (c->obs_clip)
                for (int f = 1; f < 5; f++) r[f] = clipd(r[f], -1, 1);
        }
    }
    for (int k = 0; k < K * F; k++) obs[k] = (float)rows[k];
    free(rows);
    free(cand);
    free(key);
}

/* envs/common/observation.py:354-484 OccupancyGridObservation.observe with the defaults of
 * :282-284 (features presence, vx, vy, on_road; 11x11 cells of 5 m; world axes; relative) */
static void observe_occupancy(const World *w, float *obs) {
    const NetState *s = w->s;
    const NetGraph *g = w->g;
    const int ego = ego_index(w);
    const int NX = 11, NY = 11;
    const double lo = -5.5 * 5, step = 5;
    double grid[4][11][11];
    for (int l = 0; l < 4; l++)
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < NY; j++) grid[l][i][j] = NAN;
    double ex = s->x[ego], ey = s->y[ego];
    double evx = s->speed[ego] * cos(s->heading[ego]), evy = s->speed[ego] * sin(s->heading[ego]);
    /* vehicles in REVERSED list order (df[::-1]): the lowest index owns a shared cell */
    for (int v = w->V - 1; v >= 0; v--) {
        double x = s->x[v] - ex, y = s->y[v] - ey;
        double vx = s->speed[v] * cos(s->heading[v]) - evx, vy = s->speed[v] * sin(s->heading[v]) - evy;
        vx = lmap(vx, -2 * MAX_SPEED, 2 * MAX_SPEED, -1, 1);
        vy = lmap(vy, -2 * MAX_SPEED, 2 * MAX_SPEED, -1, 1);
        int ci = (int)floor((x - lo) / step), cj = (int)floor((y - lo) / step);
        if (0 <= ci && ci < NX && 0 <= cj && cj < NY) {
            grid[0][ci][cj] = 1;
            grid[1][ci][cj] = vx;
            grid[2][ci][cj] = vy;
        }
    }
    /* fill_road_layer_by_lanes :446-484 */
    for (int l = 0; l < g->n_lanes; l++) {
        const NetLane *L = &g->lanes[l];
        double origin = lane_s(L, ex, ey);
        /* np.arange(origin - 100, origin + 100, 5): ceil((stop - start) / step) points start + k*step */
        double start = origin - 100, stop = origin + 100;
        int n = (int)ceil((stop - start) / 5.0);
        for (int k = 0; k < n; k++) {
            double wp = clipd(start + k * 5.0, 0, L->length);
            double px, py;
            net_lane_position(L, wp, 0, &px, &py);
            px -= ex;
            py -= ey;
            int ci = (int)floor((px - lo) / step), cj = (int)floor((py - lo) / step);
            if (0 <= ci && ci < NX && 0 <= cj && cj < NY) grid[3][ci][cj] = 1;
        }
    }
    for (int l = 0; l < 4; l++)
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < NY; j++) {
                double val = grid[l][i][j];
                val = clipd(val, -1, 1); /* np.clip keeps NaN */
                if (isnan(grid[l][i][j])) val = 0; /* nan_to_num */
                obs[(l * NX + i) * NY + j] = (float)val;
            }
}

int net_obs_size(const NetCfg *c) {
    if (c->obs_type == NET_OBS_OCCUPANCY) return 4 * 11 * 11;
    if (c->obs_type == NET_OBS_KINEMATICS && c->obs_n_feat > 0) return c->obs_vehicles_count * c->obs_n_feat;
    if (c->obs_type == NET_OBS_KINEMATICS) return c->obs_vehicles_count * (c->obs_features == 7 ? 7 : 5);
    if (c->obs_type == NET_OBS_TTC) return 3 * 3 * (int)(c->ttc_horizon / (1.0 / c->policy_frequency));
    return c->obs_vehicles_count * 5;
}

void net_observe(const NetGraph *g, const NetCfg *c, const NetState *s, float *obs) {
    World w;
    w.g = g;
    w.c = c;
    w.s = (NetState *)s;
    w.V = world_count(c, s);
    if (c->obs_type == NET_OBS_OCCUPANCY)
        observe_occupancy(&w, obs);
    else if (c->obs_type == NET_OBS_TTC)
        observe_ttc(&w, obs);
    else
        observe_kinematics(&w, obs);
}

/* envs/roundabout_env.py:44-71 */
static void reward_done(const World *w, int action, double *reward, int32_t *terminated,
                        int32_t *truncated) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const int ego = 0;
    const NetLane *L = LANE(w, s->lane[ego]);
    double es, elat;
    net_lane_local(L, s->x[ego], s->y[ego], &es, &elat);
    int on_road = lane_on_lane(L, es, elat, 0.0);
    double r = 0;
    r = r + c->collision_reward * (double)(s->crashed[ego] != 0);
    /* MDPVehicle.get_speed_index / (DEFAULT_TARGET_SPEEDS.size - 1) */
    r = r + c->high_speed_reward * ((double)s->speed_index[0] / (double)(3 - 1));
    r = r + c->lane_change_reward * (double)(action == 0 || action == 2);
    r = r + 0 * (double)on_road;
    if (c->normalize_reward) r = lmap(r, c->collision_reward, c->high_speed_reward, 0, 1);
    r *= (double)on_road;
    *reward = r;
    *terminated = s->crashed[ego] != 0;
    *truncated = s->time[0] >= c->duration;
}

/* envs/merge_env.py:39-84 */
static void reward_done_merge(const World *w, int action, double *reward, int32_t *terminated, int32_t *truncated) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const int ego = 0;
    double scaled_speed = lmap(s->speed[ego], c->reward_speed_lo, c->reward_speed_hi, 0, 1);
    double merging = 0; /* altruistic penalty: ControlledVehicles on the merging lane ("b", "c", 2) */
    for (int v = 0; v < w->V; v++)
        if (s->lane[v] == c->merge_lane && s->kind[v] != NET_KIND_OBSTACLE)
            merging = merging + (s->target_speed[v] - s->speed[v]) / s->target_speed[v];
    double r = 0;
    r = r + c->collision_reward * (double)(s->crashed[ego] != 0);
    r = r + c->right_lane_reward * ((double)LANE(w, s->lane[ego])->lane_id / 1);
    r = r + c->high_speed_reward * scaled_speed;
    r = r + c->lane_change_reward * (double)(action == 0 || action == 2);
    r = r + c->merging_speed_reward * merging;
    *reward = lmap(r, c->collision_reward + c->merging_speed_reward, c->high_speed_reward + c->right_lane_reward, 0, 1);
    *terminated = s->crashed[ego] != 0 || s->x[ego] > 370;
    *truncated = 0;
}

/* envs/two_way_env.py:35-62 */
static void reward_done_two_way(const World *w, double *reward, int32_t *terminated, int32_t *truncated) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const int ego = 0;
    const int n_side = LANE(w, s->lane[ego])->road_count; /* all_side_lanes(vehicle.lane_index) */
    double r = 0;
    r = r + c->high_speed_reward * ((double)s->speed_index[0] / (double)(c->n_target_speeds - 1));
    r = r + c->left_lane_reward *
                ((double)(n_side - 1 - LANE(w, s->target_lane[ego])->lane_id) / (double)(n_side - 1));
    *reward = r;
    *terminated = s->crashed[ego] != 0;
    *truncated = 0;
}

/* envs/u_turn_env.py:36-82 */
static void reward_done_u_turn(const World *w, double *reward, int32_t *terminated, int32_t *truncated) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const int ego = 0;
    const NetLane *L = LANE(w, s->lane[ego]);
    double es, elat;
    net_lane_local(L, s->x[ego], s->y[ego], &es, &elat);
    int on_road = lane_on_lane(L, es, elat, 0.0);
    int n1 = L->road_count - 1 > 1 ? L->road_count - 1 : 1;
    double scaled_speed = lmap(s->speed[ego], c->reward_speed_lo, c->reward_speed_hi, 0, 1);
    double r = 0;
    r = r + c->collision_reward * (double)(s->crashed[ego] != 0);
    r = r + c->left_lane_reward * ((double)L->lane_id / (double)n1);
    r = r + c->high_speed_reward * clipd(scaled_speed, 0, 1);
    r = r + 0 * (double)on_road;
    if (c->normalize_reward) r = lmap(r, c->collision_reward, c->high_speed_reward + c->left_lane_reward, 0, 1);
    r *= (double)on_road;
    *reward = r;
    *terminated = s->crashed[ego] != 0;
    *truncated = s->time[0] >= c->duration;
}

/* envs/exit_env.py:147-198: collision, goal (the TARGET lane is the exit lane), clipped speed term, target lane id;
 * normalised to [collision_reward, goal_reward] and clipped to [0, 1]; terminated on a crash, truncated at `duration` */
static void reward_done_exit(const World *w, double *reward, int32_t *terminated, int32_t *truncated) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const int ego = 0;
    const int tl = s->target_lane[ego];
    const int success = tl == c->exit_lane_a || tl == c->exit_lane_b;
    double scaled_speed = lmap(s->speed[ego], c->reward_speed_lo, c->reward_speed_hi, 0, 1);
    double r = 0;
    r = r + c->collision_reward * (double)(s->crashed[ego] != 0);
    r = r + c->goal_reward * (double)success;
    r = r + c->high_speed_reward * clipd(scaled_speed, 0, 1);
    r = r + c->right_lane_reward * (double)LANE(w, tl)->lane_id;
    if (c->normalize_reward) {
        r = lmap(r, c->collision_reward, c->goal_reward, 0, 1);
        r = clipd(r, 0, 1);
    }
    *reward = r;
    *terminated = s->crashed[ego] != 0;
    *truncated = s->time[0] >= c->duration;
}

/* ------------------------------------------------------------------ road/regulation.py */

/* road/road.py:323-362 position_heading_along_route(route, longitudinal, 0, current_lane_index) */
static void position_heading_along_route(const World *w, int v, double longitudinal, double *px,
                                         double *py, double *heading) {
    const NetGraph *g = w->g;
    const NetState *s = w->s;
    const int32_t *route = s->route + (size_t)v * NET_MAX_ROUTE;
    int rlen = s->route_len[v];
    int cur = s->lane[v];
    int own[1];
    if (rlen == 0) { /* `self.route or [self.lane_index]` */
        own[0] = NET_ROUTE(g->lanes[cur].from_node, g->lanes[cur].to_node, g->lanes[cur].lane_id);
        route = own;
        rlen = 1;
    }
    int k = 0;
    for (;;) {
        int first = road_first(g, RT_FROM(route[k]), RT_TO(route[k]));
        int id = RT_ID(route[k]);
        if (id < 0) id = g->lanes[cur].lane_id; /* always < len(graph[current road]) */
        const NetLane *L = &g->lanes[first + id];
        if (k < rlen - 1 && longitudinal > L->length) {
            longitudinal -= L->length;
            k++;
            continue;
        }
        net_lane_position(L, longitudinal, 0, px, py);
        *heading = net_lane_heading_at(L, longitudinal);
        return;
    }
}

/* Vehicle.predict_trajectory_constant_speed (vehicle/kinematics.py:179-198) of a NON-controlled-class vehicle (the
 * ContinuousAction ego): a deep copy acts {acceleration 0, steering = its current steering} ("constant_steering") and
 * steps through the 11 horizon points with dt = 0.25 (Vehicle.step, or BicycleVehicle.step when dynamical). */
static void predict_plain_vehicle(const World *w, int v, double px[11], double py[11], double ph[11]) {
    const NetState *s = w->s;
    double steering = w->act_steer[v], acceleration = 0.0;
    double st[6] = {s->x[v], s->y[v], s->heading[v], s->speed[v], s->lat_speed ? s->lat_speed[v] : 0.0,
                    s->yaw_rate ? s->yaw_rate[v] : 0.0};
    int crashed = s->crashed[v], has_impact = s->has_impact[v];
    for (int k = 0; k < 11; k++) {
        if (w->c->dynamical) {
            bicycle_advance(st, crashed, &steering, &acceleration, 0.25);
        } else {
            kinematic_advance(st, crashed, &steering, &acceleration, 0.25);
            if (has_impact) { /* kinematics.py:147-150 (applied after the position update) */
                st[0] += s->impact_x[v];
                st[1] += s->impact_y[v];
                crashed = 1;
                has_impact = 0;
            }
        }
        px[k] = st[0];
        py[k] = st[1];
        ph[k] = st[2];
    }
}

/* regulation.py:85-111 is_conflict_possible (horizon 3, step 0.25) */
static int is_conflict_possible(const World *w, int v1, int v2) {
    const NetState *s = w->s;
    double s1 = lane_s(LANE(w, s->lane[v1]), s->x[v1], s->y[v1]);
    double s2 = lane_s(LANE(w, s->lane[v2]), s->x[v2], s->y[v2]);
    double q1x[11], q1y[11], q1h[11], q2x[11], q2y[11], q2h[11];
    const int plain1 = s->kind[v1] == NET_KIND_VEHICLE, plain2 = s->kind[v2] == NET_KIND_VEHICLE;
    if (plain1) predict_plain_vehicle(w, v1, q1x, q1y, q1h);
    if (plain2) predict_plain_vehicle(w, v2, q2x, q2y, q2h);
    for (int k = 1; k < 12; k++) {
        double t = 0.25 * k; /* np.arange(0.25, 3, 0.25) */
        double p1x, p1y, h1, p2x, p2y, h2;
        if (plain1) {
            p1x = q1x[k - 1], p1y = q1y[k - 1], h1 = q1h[k - 1];
        } else {
            position_heading_along_route(w, v1, s1 + s->speed[v1] * t, &p1x, &p1y, &h1);
        }
        if (plain2) {
            p2x = q2x[k - 1], p2y = q2y[k - 1], h2 = q2h[k - 1];
        } else {
            position_heading_along_route(w, v2, s2 + s->speed[v2] * t, &p2x, &p2y, &h2);
        }
        if (norm2(p2x - p1x, p2y - p1y) > VEH_LENGTH) continue;
        if (orc_rotated_rectangles_intersect(p1x, p1y, 1.5 * VEH_LENGTH, 0.9 * VEH_WIDTH, h1, p2x, p2y,
                                             1.5 * VEH_LENGTH, 0.9 * VEH_WIDTH, h2))
            return 1;
    }
    return 0;
}

/* regulation.py:42-83 enforce_road_rules + respect_priorities (YIELD_DURATION = 0) */
static void enforce_road_rules(World *w) {
    NetState *s = w->s;
    for (int v = 0; v < w->V; v++)
        if (s->is_yielding[v]) {
            s->target_speed[v] = LANE(w, s->lane[v])->speed_limit;
            s->is_yielding[v] = 0;
        }
    for (int i = 0; i < w->V - 1; i++)
        for (int j = i + 1; j < w->V; j++) {
            if (!is_conflict_possible(w, i, j)) continue;
            int p1 = LANE(w, s->lane[i])->priority, p2 = LANE(w, s->lane[j])->priority;
            int y;
            if (p1 > p2)
                y = j;
            else if (p1 < p2)
                y = i;
            else {
                double f12 = dot2(cos(s->heading[i]), sin(s->heading[i]), s->x[j] - s->x[i], s->y[j] - s->y[i]);
                double f21 = dot2(cos(s->heading[j]), sin(s->heading[j]), s->x[i] - s->x[j], s->y[i] - s->y[j]);
                y = f12 > f21 ? i : j;
            }
            if (s->kind[y] == NET_KIND_IDM) { /* ControlledVehicle and not MDPVehicle */
                s->target_speed[y] = 0;
                s->is_yielding[y] = 1;
            }
        }
}

/* regulation.py:36-40 RegulatedRoad.step prologue */
static void regulated_pre_step(World *w, double dt) {
    w->s->road_steps[0] += 1;
    if (w->s->road_steps[0] % (int)(1 / dt / 2) == 0) enforce_road_rules(w);
}

/* envs/intersection_env.py:368-373 */
int net_has_arrived(const NetGraph *g, const NetState *s, int v) {
    const NetLane *L = &g->lanes[s->lane[v]];
    return L->exit_lane && lane_s(L, s->x[v], s->y[v]) >= 25;
}

/* envs/intersection_env.py:93-117 _agent_reward / _agent_rewards of one controlled vehicle */
static double agent_reward_intersection(const World *w, int ego, int *arrived_out, int *on_road_out) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const NetLane *L = LANE(w, s->lane[ego]);
    double es, elat;
    net_lane_local(L, s->x[ego], s->y[ego], &es, &elat);
    int on_road = lane_on_lane(L, es, elat, 0.0);
    int arrived = net_has_arrived(w->g, s, ego);
    double scaled_speed = lmap(s->speed[ego], c->reward_speed_lo, c->reward_speed_hi, 0, 1);
    double r = 0;
    r = r + c->collision_reward * (double)(s->crashed[ego] != 0);
    r = r + c->high_speed_reward * clipd(scaled_speed, 0, 1);
    r = r + c->arrived_reward * (double)arrived;
    r = r + 0 * (double)on_road;
    if (arrived) r = c->arrived_reward;
    r *= (double)on_road;
    if (c->normalize_reward) r = lmap(r, c->collision_reward, c->arrived_reward, 0, 1);
    *arrived_out = arrived;
    *on_road_out = on_road;
    return r;
}

/* envs/intersection_env.py:79-117 (single controlled vehicle) */
static void reward_done_intersection(const World *w, double *reward, int32_t *terminated,
                                     int32_t *truncated) {
    const NetCfg *c = w->c;
    const NetState *s = w->s;
    const int ego = ego_index(w);
    const NetLane *L = LANE(w, s->lane[ego]);
    double es, elat;
    net_lane_local(L, s->x[ego], s->y[ego], &es, &elat);
    int on_road = lane_on_lane(L, es, elat, 0.0);
    int arrived = net_has_arrived(w->g, s, ego);
    double scaled_speed = lmap(s->speed[ego], c->reward_speed_lo, c->reward_speed_hi, 0, 1);
    double r = 0;
    r = r + c->collision_reward * (double)(s->crashed[ego] != 0);
    r = r + c->high_speed_reward * clipd(scaled_speed, 0, 1);
    r = r + c->arrived_reward * (double)arrived;
    r = r + 0 * (double)on_road;
    if (arrived) r = c->arrived_reward;
    r *= (double)on_road;
    if (c->normalize_reward) r = lmap(r, c->collision_reward, c->arrived_reward, 0, 1);
    *reward = r / 1; /* mean over the (single) controlled vehicle */
    *terminated = s->crashed[ego] != 0 || arrived || (c->offroad_terminal && !on_road);
    *truncated = s->time[0] >= c->duration;
}

/* envs/common/action.py:136-162 ContinuousAction.get_action / act: the Box action is float32 and NEP 50 keeps
 * utils.lmap (utils.py:31-33) in float32; the vehicle's action dict then holds the widened values */
static void continuous_act(World *w, int v, const float *a) {
    const NetCfg *c = w->c;
    float a0 = a[0], a1 = a[1];
    if (c->act_clip) {
        a0 = fminf(fmaxf(a0, -1.0f), 1.0f);
        a1 = fminf(fmaxf(a1, -1.0f), 1.0f);
    }
    float acc = (float)c->acc_lo + (a0 - (-1.0f)) * (float)(c->acc_hi - c->acc_lo) / 2.0f;
    float steer = (float)c->steer_lo + (a1 - (-1.0f)) * (float)(c->steer_hi - c->steer_lo) / 2.0f;
    w->act_accel[v] = (double)acc;
    w->act_steer[v] = (double)steer;
}

static void net_step_any(const NetGraph *g, const NetCfg *c, NetState *s, int action, const float *action_f,
                         float *obs, double *reward, int32_t *terminated, int32_t *truncated);
void net_step(const NetGraph *g, const NetCfg *c, NetState *s, int action, float *obs, double *reward,
              int32_t *terminated, int32_t *truncated) {
    net_step_any(g, c, s, action, NULL, obs, reward, terminated, truncated);
}
void net_step_continuous(const NetGraph *g, const NetCfg *c, NetState *s, const float *action, float *obs,
                         double *reward, int32_t *terminated, int32_t *truncated) {
    net_step_any(g, c, s, -1, action, obs, reward, terminated, truncated);
}

/* envs/common/abstract.py:259-317 */
static void net_step_any(const NetGraph *g, const NetCfg *c, NetState *s, int action, const float *action_f,
                         float *obs, double *reward, int32_t *terminated, int32_t *truncated) {
    World w;
    w.g = g;
    w.c = c;
    w.s = s;
    w.V = world_count(c, s);
    double *act_buf = (double *)calloc(2 * (size_t)w.V, sizeof(double));
    w.act_steer = act_buf;
    w.act_accel = act_buf + w.V;
    int frames = c->simulation_frequency / c->policy_frequency;
    double dt = 1.0 / c->simulation_frequency;
    s->time[0] += 1.0 / c->policy_frequency;
    /* intersection: the ego is the LAST vehicle of the list (appended after the traffic) */
    int ego = ego_index(&w);
    int label = action;
    if (c->action_mode == 1) label = action == 0 ? 4 : (action == 2 ? 3 : 1); /* SLOWER / IDLE / FASTER */
    for (int frame = 0; frame < frames; frame++) {
        if (frame == 0) {
            if (action_f)
                continuous_act(&w, ego, action_f);
            else
                mdp_act(&w, ego, label);
        }
        road_act(&w);
        if (c->regulated) regulated_pre_step(&w, dt);
        road_step(&w, dt);
    }
    if (obs) net_observe(g, c, s, obs);
    if (c->reward_type == 1)
        reward_done_intersection(&w, reward, terminated, truncated);
    else if (c->reward_type == 2)
        reward_done_merge(&w, action, reward, terminated, truncated);
    else if (c->reward_type == 3)
        reward_done_two_way(&w, reward, terminated, truncated);
    else if (c->reward_type == 4)
        reward_done_u_turn(&w, reward, terminated, truncated);
    else if (c->reward_type == 5)
        reward_done_exit(&w, reward, terminated, truncated);
    else
        reward_done(&w, action, reward, terminated, truncated);
    free(act_buf);
}

/* The same step with several controlled vehicles: MultiAgentAction.act (envs/common/action.py:301-333) applies
 * actions[k] to controlled vehicle k in order; MultiAgentObservation (observation.py:588-604) stacks one
 * observation per agent; _reward is the mean of the agents' rewards, the episode terminates when ANY agent crashed
 * or ALL arrived (intersection_env.py:79-134); info carries the per-agent rewards and terminal flags. */
void net_step_agents(const NetGraph *g, const NetCfg *c, NetState *s, const int32_t *actions, int n_agents,
                     float *obs, double *reward, int32_t *terminated, int32_t *truncated, double *agents_reward,
                     int32_t *agents_terminated) {
    World w;
    w.g = g;
    w.c = c;
    w.s = s;
    w.V = world_count(c, s);
    double *act_buf = (double *)calloc(2 * (size_t)w.V, sizeof(double));
    w.act_steer = act_buf;
    w.act_accel = act_buf + w.V;
    int frames = c->simulation_frequency / c->policy_frequency;
    double dt = 1.0 / c->simulation_frequency;
    s->time[0] += 1.0 / c->policy_frequency;
    for (int frame = 0; frame < frames; frame++) {
        if (frame == 0)
            for (int k = 0; k < n_agents; k++) {
                int v = agent_vehicle(&w, k);
                int label = actions[k];
                if (c->action_mode == 1) label = actions[k] == 0 ? 4 : (actions[k] == 2 ? 3 : 1);
                if (v >= 0) mdp_act_agent(&w, v, label, k);
            }
        road_act(&w);
        if (c->regulated) regulated_pre_step(&w, dt);
        road_step(&w, dt);
    }
    const int per = net_obs_size(c);
    double sum = 0;
    int any_crashed = 0, all_arrived = 1, first_on_road = 1;
    for (int k = 0; k < n_agents; k++) {
        int v = agent_vehicle(&w, k);
        if (obs) observe_kinematics_from(&w, v, obs + (size_t)k * per);
        int arrived, on_road;
        double r = agent_reward_intersection(&w, v, &arrived, &on_road);
        sum += r;
        any_crashed |= s->crashed[v] != 0;
        all_arrived &= arrived;
        if (k == 0) first_on_road = on_road;
        if (agents_reward) agents_reward[k] = r;
        if (agents_terminated) agents_terminated[k] = s->crashed[v] != 0 || arrived;
    }
    *reward = sum / n_agents;
    *terminated = any_crashed || all_arrived || (c->offroad_terminal && !first_on_road);
    *truncated = s->time[0] >= c->duration;
    free(act_buf);
}

void net_observe_agents(const NetGraph *g, const NetCfg *c, const NetState *s, int n_agents, float *obs) {
    World w;
    w.g = g;
    w.c = c;
    w.s = (NetState *)s;
    w.V = world_count(c, s);
    for (int k = 0; k < n_agents; k++) observe_kinematics_from(&w, agent_vehicle(&w, k), obs + (size_t)k * net_obs_size(c));
}

void net_substeps(const NetGraph *g, const NetCfg *c, NetState *s, int substeps) {
    World w;
    w.g = g;
    w.c = c;
    w.s = s;
    w.V = world_count(c, s);
    double *act_buf = (double *)calloc(2 * (size_t)(w.V > 0 ? w.V : 1), sizeof(double));
    w.act_steer = act_buf;
    w.act_accel = act_buf + w.V;
    double dt = 1.0 / c->simulation_frequency;
    for (int k = 0; k < substeps; k++) {
        road_act(&w);
        if (c->regulated) regulated_pre_step(&w, dt);
        road_step(&w, dt);
    }
    free(act_buf);
}


/* ====================================================================== observation plugins, general form
 * Vehicle.to_dict (vehicle/kinematics.py:237-261) relative to `origin` (< 0: absolute); road objects
 * (vehicle/objects.py:141-159) have no heading / offsets columns (NaN in the DataFrame). */
static double vehicle_feature(const NetGraph *g, const NetState *s, int v, int origin, int feat, int observe_intentions) {
    const int is_object = s->kind[v] == NET_KIND_OBSTACLE;
    const double ch = cos(s->heading[v]), sh = sin(s->heading[v]);
    switch (feat) {
        case NET_FEAT_PRESENCE: return 1.0;
        case NET_FEAT_X: return origin >= 0 ? s->x[v] - s->x[origin] : s->x[v];
        case NET_FEAT_Y: return origin >= 0 ? s->y[v] - s->y[origin] : s->y[v];
        case NET_FEAT_VX: {
            double vx = is_object ? 0.0 : s->speed[v] * ch;
            return origin >= 0 ? vx - s->speed[origin] * cos(s->heading[origin]) : vx;
        }
        case NET_FEAT_VY: {
            double vy = is_object ? 0.0 : s->speed[v] * sh;
            return origin >= 0 ? vy - s->speed[origin] * sin(s->heading[origin]) : vy;
        }
        case NET_FEAT_HEADING: return is_object ? NAN : s->heading[v];
        case NET_FEAT_COS_H: return ch;
        case NET_FEAT_SIN_H: return sh;
        case NET_FEAT_COS_D:
        case NET_FEAT_SIN_D: {
            /* destination (:203-215): end of the last route lane (lane id None -> 0); no route: the position */
            int rl = s->route_len ? s->route_len[v] : 0;
            if (!observe_intentions || is_object || rl == 0) return 0.0;
            int e = s->route[(size_t)v * NET_MAX_ROUTE + rl - 1];
            int first = road_first(g, RT_FROM(e), RT_TO(e));
            int id = RT_ID(e) < 0 ? 0 : RT_ID(e);
            const NetLane *L = &g->lanes[first + id];
            double dx, dy;
            net_lane_position(L, L->length, 0, &dx, &dy);
            dx -= s->x[v];
            dy -= s->y[v];
            if (dx == 0.0 && dy == 0.0) return 0.0;
            double n = sqrt(dx * dx + dy * dy); /* np.linalg.norm of a 2-vector */
            return feat == NET_FEAT_COS_D ? dx / n : dy / n;
        }
        case NET_FEAT_LONG_OFF:
        case NET_FEAT_LAT_OFF:
        case NET_FEAT_ANG_OFF: {
            if (is_object) return NAN;
            const NetLane *L = &g->lanes[s->lane[v]];
            double lon, lat;
            net_lane_local(L, s->x[v], s->y[v], &lon, &lat);
            if (feat == NET_FEAT_LONG_OFF) return lon;
            if (feat == NET_FEAT_LAT_OFF) return lat;
            return orc_wrap_to_pi(s->heading[v] - net_lane_heading_at(L, lon)); /* lane.local_angle :145-147 */
        }
        default: return NAN;
    }
}

/* OccupancyGridObservation.pos_to_index (:422-444); the position is relative to the observer */
static void grid_index(const NetGridCfg *gc, double px, double py, double ce, double se, int *ci, int *cj) {
    if (gc->align_to_vehicle_axes) { /* [[c, s], [-s, c]] @ position */
        double rx = ce * px + se * py, ry = -se * px + ce * py;
        px = rx;
        py = ry;
    }
    *ci = (int)floor((px - gc->grid_lo[0]) / gc->grid_step[0]);
    *cj = (int)floor((py - gc->grid_lo[1]) / gc->grid_step[1]);
}

/* OccupancyGridObservation.observe (:354-420) with every constructor option */
void net_observe_grid(const NetGraph *g, const NetState *s, int V, int ego, const NetGridCfg *gc, float *obs) {
    const int NX = gc->shape[0], NY = gc->shape[1], F = gc->n_features;
    const size_t cells = (size_t)NX * NY;
    double *grid = (double *)malloc(sizeof(double) * cells * F);
    for (size_t k = 0; k < cells * F; k++) grid[k] = NAN;
    const double ex = s->x[ego], ey = s->y[ego];
    const double ce = cos(s->heading[ego]), se = sin(s->heading[ego]);
    for (int layer = 0; layer < F; layer++) {
        const int feat = gc->features[layer];
        if (feat == NET_FEAT_ON_ROAD) {
            /* fill_road_layer_by_lanes (:466-499) */
            const double spacing = fmin(gc->grid_step[0], gc->grid_step[1]);
            for (int l = 0; l < g->n_lanes; l++) {
                const NetLane *L = &g->lanes[l];
                double origin = lane_s(L, ex, ey);
                double start = origin - 100, stop = origin + 100;
                int n = (int)ceil((stop - start) / spacing); /* np.arange length */
                for (int k = 0; k < n; k++) {
                    double wp = clipd(start + k * spacing, 0, L->length);
                    double px, py;
                    net_lane_position(L, wp, 0, &px, &py);
                    int ci, cj;
                    grid_index(gc, px - ex, py - ey, ce, se, &ci, &cj);
                    if (0 <= ci && ci < NX && 0 <= cj && cj < NY) grid[(size_t)layer * cells + (size_t)ci * NY + cj] = 1;
                }
            }
            continue;
        }
        if (feat == NET_FEAT_UNKNOWN) continue;
        for (int v = V - 1; v >= 0; v--) { /* df[::-1]: the lowest index is written last */
            if (s->kind[v] == NET_KIND_OBSTACLE) continue; /* road.vehicles only */
            double x = vehicle_feature(g, s, v, ego, NET_FEAT_X, 1), y = vehicle_feature(g, s, v, ego, NET_FEAT_Y, 1);
            /* normalize() maps x / y when they have a range; the cell index un-maps them (:383-400) */
            if (gc->x_ranged) x = lmap(lmap(x, gc->x_lo, gc->x_hi, -1, 1), -1, 1, gc->x_lo, gc->x_hi);
            if (gc->y_ranged) y = lmap(lmap(y, gc->y_lo, gc->y_hi, -1, 1), -1, 1, gc->y_lo, gc->y_hi);
            int ci, cj;
            grid_index(gc, x, y, ce, se, &ci, &cj);
            if (!(0 <= ci && ci < NX && 0 <= cj && cj < NY)) continue;
            double val = vehicle_feature(g, s, v, ego, feat, gc->observe_intentions);
            if (gc->ranged[layer]) val = lmap(val, gc->range_lo[layer], gc->range_hi[layer], -1, 1);
            grid[(size_t)layer * cells + (size_t)ci * NY + cj] = val;
        }
    }
    for (size_t k = 0; k < cells * F; k++) {
        double val = grid[k];
        if (isnan(val)) { /* np.clip keeps NaN; astype(uint8) of NaN and nan_to_num both give 0 */
            obs[k] = 0.0f;
            continue;
        }
        if (gc->clip) val = clipd(val, -1, 1);
        if (gc->as_image) val = (double)(unsigned char)(long)((clipd(val, -1, 1) + 1) / 2 * 255); /* .astype(np.uint8) */
        obs[k] = (float)val;
    }
    free(grid);
}

/* utils.distance_to_rect (utils.py:388-416): ray [r, q] against the rectangle (a, b, c, d) */
static double lidar_distance_to_rect(const double r[2], const double q[2], double corners[4][2]) {
    const double *a = corners[0], *b = corners[1], *d = corners[3];
    double ux = b[0] - a[0], uy = b[1] - a[1], vx = d[0] - a[0], vy = d[1] - a[1];
    double un = norm2(ux, uy), vn = norm2(vx, vy);
    ux /= un;
    uy /= un;
    vx /= vn;
    vy /= vn;
    double rqu = dot2(q[0] - r[0], q[1] - r[1], ux, uy), rqv = dot2(q[0] - r[0], q[1] - r[1], vx, vy);
    double i1[2] = {dot2(a[0] - r[0], a[1] - r[1], ux, uy) / rqu, dot2(b[0] - r[0], b[1] - r[1], ux, uy) / rqu};
    double i2[2] = {dot2(a[0] - r[0], a[1] - r[1], vx, vy) / rqv, dot2(d[0] - r[0], d[1] - r[1], vx, vy) / rqv};
    if (!(rqu >= 0)) {
        double t = i1[0];
        i1[0] = i1[1];
        i1[1] = t;
    }
    if (!(rqv >= 0)) {
        double t = i2[0];
        i2[0] = i2[1];
        i2[1] = t;
    }
    if (interval_distance(i1[0], i1[1], i2[0], i2[1]) <= 0 && interval_distance(0, 1, i1[0], i1[1]) <= 0 &&
        interval_distance(0, 1, i2[0], i2[1]) <= 0)
        return fmax(i1[0], i2[0]) * norm2(q[0] - r[0], q[1] - r[1]);
    return INFINITY;
}

/* LidarObservation.trace / observe (observation.py:703-769).  The grid is float32: every store rounds, and the
 * `<=` comparisons read the rounded values back. */
void net_observe_lidar(const NetState *s, int V, int ego, int cells, double maximum_range, int normalize, float *obs) {
    const double angle = 2 * M_PI / cells;
    float *grid = obs;
    for (int k = 0; k < cells; k++) grid[2 * k] = grid[2 * k + 1] = (float)maximum_range;
    const double ox = s->x[ego], oy = s->y[ego];
    const double ovx = s->speed[ego] * cos(s->heading[ego]), ovy = s->speed[ego] * sin(s->heading[ego]);
#define ANGLE_TO_INDEX(a) ((int)py_mod(floor((a) / angle), (double)cells))
    for (int o = 0; o < V; o++) { /* road.vehicles + road.objects: objects sit after the vehicles */
        if (o == ego) continue;
        const int is_object = s->kind[o] == NET_KIND_OBSTACLE;
        const double len = is_object ? 2.0 : VEH_LENGTH, wid = 2.0;
        const double ch = cos(s->heading[o]), sh = sin(s->heading[o]);
        const double vx = is_object ? 0.0 : s->speed[o] * ch, vy = is_object ? 0.0 : s->speed[o] * sh;
        double center_distance = norm2(s->x[o] - ox, s->y[o] - oy);
        if (center_distance > maximum_range) continue;
        double center_angle = atan2(s->y[o] - oy, s->x[o] - ox) + angle / 2;
        int center_index = ANGLE_TO_INDEX(center_angle);
        double distance = center_distance - wid / 2;
        if (distance <= grid[2 * center_index]) {
            double dx = cos(center_index * angle), dy = sin(center_index * angle);
            grid[2 * center_index] = (float)distance;
            grid[2 * center_index + 1] = (float)dot2(vx - ovx, vy - ovy, dx, dy);
        }
        /* utils.rect_corners (utils.py:128-157): rotation @ corner + centre */
        const double cl[4][2] = {{-len / 2, -wid / 2}, {-len / 2, wid / 2}, {len / 2, wid / 2}, {len / 2, -wid / 2}};
        double corners[4][2], amin = 0, amax = 0;
        for (int k = 0; k < 4; k++) {
            corners[k][0] = (ch * cl[k][0] + (-sh) * cl[k][1]) + s->x[o];
            corners[k][1] = (sh * cl[k][0] + ch * cl[k][1]) + s->y[o];
            double a = atan2(corners[k][1] - oy, corners[k][0] - ox) + angle / 2;
            if (k == 0 || a < amin) amin = a;
            if (k == 0 || a > amax) amax = a;
        }
        if (amin < -M_PI / 2 && M_PI / 2 < amax) { /* corners wrap around +pi */
            double t = amin;
            amin = amax;
            amax = t + 2 * M_PI;
        }
        int start = ANGLE_TO_INDEX(amin), end = ANGLE_TO_INDEX(amax);
        int n_idx = start < end ? end - start + 1 : (cells - start) + (end + 1);
        for (int q = 0; q < n_idx; q++) {
            int index = start < end ? start + q : (q < cells - start ? start + q : q - (cells - start));
            double dx = cos(index * angle), dy = sin(index * angle);
            double r[2] = {ox, oy}, qq[2] = {ox + maximum_range * dx, oy + maximum_range * dy};
            double dist = lidar_distance_to_rect(r, qq, corners);
            if (dist <= grid[2 * index]) {
                grid[2 * index] = (float)dist;
                grid[2 * index + 1] = (float)dot2(vx - ovx, vy - ovy, dx, dy);
            }
        }
    }
#undef ANGLE_TO_INDEX
    if (normalize) /* obs /= maximum_range on the float32 array */
        for (int k = 0; k < 2 * cells; k++) grid[k] = grid[k] / (float)maximum_range;
}
