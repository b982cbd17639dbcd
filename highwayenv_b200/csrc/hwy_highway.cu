// hwy_highway.cu — sm_100a kernels + C ABI for the straight-highway family
// (highway-v0 / highway-fast-v0) of the batched HighwayEnv hot path.
//
// One (env, vehicle) pair per thread; TPE threads per env (32/64/128, the next power of
// two >= n_vehicles).  The whole AbstractEnv.step — all substeps of Road.act/Road.step,
// then observe/reward/termination — runs in ONE kernel: the SoA state makes one HBM round
// trip per env-step, neighbour data is staged in shared memory once per substep, and the
// O(V^2) searches (IDM front/rear vehicle, MOBIL, abort scan, collision sweep) broadcast
// from shared memory.  The reference's sequential semantics (Gauss-Seidel target-lane
// updates in Road.act, last-writer-wins impacts in Road.step) are reproduced with an
// ordered bit-mask fix-up, see DESIGN.md.
//
// Reference paths are relative to /root/reference/highway_env.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/hwyb200.h"
#include "hwy_math.cuh"

namespace hwy {

typedef unsigned long long u64;

// ------------------------------------------------------------------ lane geometry
// road/lane.py:205-209 StraightLane.local_coordinates
__device__ __forceinline__ void lane_local(const HwyStraightLane& L, double x, double y, double& s,
                                           double& lat) {
    double ddx = x - L.start_x, ddy = y - L.start_y;
    s = dot2(ddx, ddy, L.dir_x, L.dir_y);
    lat = dot2(ddx, ddy, L.lat_x, L.lat_y);
}
__device__ __forceinline__ double lane_s(const HwyStraightLane& L, double x, double y) {
    return dot2(x - L.start_x, y - L.start_y, L.dir_x, L.dir_y);
}
// road/lane.py:80-102 on_lane
__device__ __forceinline__ bool lane_on(const HwyStraightLane& L, double s, double lat, double margin) {
    return fabs(lat) <= L.width / 2 + margin && -kLaneVehLength <= s && s < L.length + kLaneVehLength;
}
// road/lane.py:104-118 is_reachable_from (forbidden is False on the highway)
__device__ __forceinline__ bool lane_reachable(const HwyStraightLane& L, double x, double y) {
    double s, lat;
    lane_local(L, x, y, s, lat);
    return fabs(lat) <= 2 * L.width && 0 <= s && s < L.length + kLaneVehLength;
}
// road/road.py:55-71 get_closest_lane_index with lane.py:132-143 distance_with_heading
__device__ __forceinline__ int closest_lane(const HwyHighwayParams& P, double x, double y, double h) {
    int best = 0;
    double bd = 0;
    for (int l = 0; l < P.lanes_count; ++l) {
        const HwyStraightLane& L = P.lanes[l];
        double s, r;
        lane_local(L, x, y, s, r);
        double angle = fabs(wrap_to_pi(h - L.heading));
        double d = fabs(r) + fmax(s - L.length, 0.0) + fmax(0.0 - s, 0.0) + 1.0 * angle;
        if (l == 0 || d < bd) {
            bd = d;
            best = l;
        }
    }
    return best;
}

// ------------------------------------------------------------------ shared staging
template <int TPE>
struct EnvShared {
    static constexpr int NW = (TPE + 63) / 64;
    double x[TPE], y[TPE], c[TPE], s[TPE], v[TPE], ts[TPE];
    double key[TPE];  // observation sort keys
    u64 geo[TPE][NW];             // abort-scan geometric candidates of mid-change vehicles
    u64 tm[HWY_MAX_LANES][NW];    // vehicles whose current target lane is l
    u64 lane_ne[HWY_MAX_LANES][NW];  // vehicles whose lane_index != l
    u64 mid[NW];                  // active mid-change IDM vehicles (lane != target)
    u64 changed[NW];              // IDM vehicles whose MOBIL decision changed the target
    u64 aborted[NW];
    unsigned char lane[TPE], tgt[TPE], tgt1[TPE], flags[TPE];  // flags: 1 check_collisions, 2 controlled
};

template <int TPE>
__device__ __forceinline__ void env_sync() {
    if (TPE <= 32)
        __syncwarp();
    else
        __syncthreads();
}

// road/road.py:483-547 neighbour_vehicles (same-segment search).  Ties: front `<=` keeps the
// later index, rear `>` keeps the earlier one.
template <int TPE>
__device__ __forceinline__ void neighbours(const EnvShared<TPE>& sm, const HwyStraightLane& L, int V,
                                           int self, int& front, int& rear) {
    double s = lane_s(L, sm.x[self], sm.y[self]);
    double s_front = 0, s_rear = 0;
    front = -1;
    rear = -1;
    for (int v = 0; v < V; ++v) {
        if (v == self) continue;
        double s_v, lat_v;
        lane_local(L, sm.x[v], sm.y[v], s_v, lat_v);
        if (!lane_on(L, s_v, lat_v, 1.0)) continue;
        if (s <= s_v && (front < 0 || s_v <= s_front)) {
            s_front = s_v;
            front = v;
        }
        if (s_v < s && (rear < 0 || s_v > s_rear)) {
            s_rear = s_v;
            rear = v;
        }
    }
}

// vehicle/behavior.py:192-217 desired_gap(ego, front), projected
template <int TPE>
__device__ __forceinline__ double desired_gap(const HwyHighwayParams& P, const EnvShared<TPE>& sm,
                                              int ego, int front) {
    double ab = -P.comfort_acc_max * P.comfort_acc_min;
    double dvx = sm.v[ego] * sm.c[ego] - sm.v[front] * sm.c[front];
    double dvy = sm.v[ego] * sm.s[ego] - sm.v[front] * sm.s[front];
    double dv = dot2(dvx, dvy, sm.c[ego], sm.s[ego]);
    return P.distance_wanted + sm.v[ego] * P.time_wanted + sm.v[ego] * dv / (2 * sqrt(ab));
}

// vehicle/behavior.py:150-190 acceleration(ego_vehicle, front_vehicle) with the CALLER's DELTA
template <int TPE>
__device__ __forceinline__ double idm_acceleration(const HwyHighwayParams& P, const EnvShared<TPE>& sm,
                                                   double delta, int ego, int front) {
    if (ego < 0) return 0.0;
    const HwyStraightLane& L = P.lanes[sm.lane[ego]];
    double ego_target_speed = clipd(sm.ts[ego], 0.0, L.speed_limit);
    double acc = P.comfort_acc_max *
                 (1 - pow(fmax(sm.v[ego], 0.0) / fabs(not_zero(ego_target_speed)), delta));
    if (front >= 0) {
        double d = lane_s(L, sm.x[front], sm.y[front]) - lane_s(L, sm.x[ego], sm.y[ego]);
        double q = desired_gap(P, sm, ego, front) / not_zero(d);
        acc -= P.comfort_acc_max * (q * q);  // np.power(q, 2)
    }
    return acc;
}

// vehicle/controller.py:145-187 steering_control on a StraightLane
__device__ __forceinline__ double steering_control(const HwyStraightLane& L, double x, double y,
                                                   double heading, double speed) {
    double lc_s, lc_lat;
    lane_local(L, x, y, lc_s, lc_lat);
    double lane_future_heading = L.heading;  // StraightLane.heading_at
    double lateral_speed_command = -kKpLateral * lc_lat;
    double heading_command = asin(clipd(lateral_speed_command / not_zero(speed), -1.0, 1.0));
    double heading_ref = lane_future_heading + clipd(heading_command, -kPi / 4, kPi / 4);
    double heading_rate_command = kKpHeading * wrap_to_pi(heading_ref - heading);
    double slip_angle =
        asin(clipd(kVehLength / 2 / not_zero(speed) * heading_rate_command, -1.0, 1.0));
    double steering_angle = atan(2 * tan(slip_angle));
    return clipd(steering_angle, -kMaxSteer, kMaxSteer);
}

// vehicle/behavior.py:265-324 mobil(lane_index); route is None on the highway
template <int TPE>
__device__ __forceinline__ bool mobil(const HwyHighwayParams& P, const EnvShared<TPE>& sm, int V, int i,
                                      double delta, int cand, int old_preceding, int old_following) {
    int new_preceding, new_following;
    neighbours(sm, P.lanes[cand], V, i, new_preceding, new_following);
    double new_following_pred_a = idm_acceleration(P, sm, delta, new_following, i);
    if (new_following_pred_a < -P.lane_change_max_braking_imposed) return false;
    double self_pred_a = idm_acceleration(P, sm, delta, i, new_preceding);
    double self_a = idm_acceleration(P, sm, delta, i, old_preceding);
    double jerk = self_pred_a - self_a;
    if (P.politeness != 0.0) {
        double new_following_a = idm_acceleration(P, sm, delta, new_following, new_preceding);
        double old_following_a = idm_acceleration(P, sm, delta, old_following, i);
        double old_following_pred_a = idm_acceleration(P, sm, delta, old_following, old_preceding);
        jerk = self_pred_a - self_a +
               P.politeness * (new_following_pred_a - new_following_a + old_following_pred_a -
                               old_following_a);
    }
    return !(jerk < P.lane_change_min_acc_gain);
}

// vehicle/objects.py:169-181 polygon()
template <int TPE>
__device__ __forceinline__ void polygon(const EnvShared<TPE>& sm, int v, double (&p)[5][2]) {
    const double hl = kVehLength / 2, hw = kVehWidth / 2;
    const double lx[4] = {-hl, -hl, +hl, +hl};
    const double ly[4] = {-hw, +hw, +hw, -hw};
    double c = sm.c[v], s = sm.s[v];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        p[k][0] = (c * lx[k] + (-s) * ly[k]) + sm.x[v];
        p[k][1] = (s * lx[k] + c * ly[k]) + sm.y[v];
    }
    p[4][0] = p[0][0];
    p[4][1] = p[0][1];
}

__device__ __forceinline__ void project_polygon(const double (&p)[5][2], double ax, double ay,
                                                double& mn, double& mx) {
    mn = mx = dot2(p[0][0], p[0][1], ax, ay);
#pragma unroll
    for (int k = 1; k < 5; ++k) {
        double pr = dot2(p[k][0], p[k][1], ax, ay);
        if (pr < mn) mn = pr;
        if (pr > mx) mx = pr;
    }
}
__device__ __forceinline__ double interval_distance(double min_a, double max_a, double min_b,
                                                    double max_b) {
    return min_a < min_b ? min_b - max_a : min_a - max_b;
}

// utils.py:196-241 are_polygons_intersecting: SAT over the 4+4 edge normals with the
// relative displacement extension; returns (intersecting, will_intersect, translation).
__device__ __noinline__ void polygons_intersecting(const double (&a)[5][2], const double (&b)[5][2],
                                                   double dax, double day, double dbx, double dby,
                                                   bool& intersecting, bool& will_intersect,
                                                   double& trx, double& try_) {
    intersecting = true;
    will_intersect = true;
    double min_distance = INFINITY;
    double tax = 0, tay = 0;
    double cax = (((a[0][0] + a[1][0]) + a[2][0]) + a[3][0]) / 4.0;
    double cay = (((a[0][1] + a[1][1]) + a[2][1]) + a[3][1]) / 4.0;
    double cbx = (((b[0][0] + b[1][0]) + b[2][0]) + b[3][0]) / 4.0;
    double cby = (((b[0][1] + b[1][1]) + b[2][1]) + b[3][1]) / 4.0;
    double dcx = cax - cbx, dcy = cay - cby;
    for (int poly = 0; poly < 2; ++poly) {
        for (int e = 0; e < 4; ++e) {
            double p1x = poly == 0 ? a[e][0] : b[e][0], p1y = poly == 0 ? a[e][1] : b[e][1];
            double p2x = poly == 0 ? a[e + 1][0] : b[e + 1][0];
            double p2y = poly == 0 ? a[e + 1][1] : b[e + 1][1];
            double nx = -p2y + p1y, ny = p2x - p1x;
            double nn = norm2(nx, ny);
            nx /= nn;
            ny /= nn;
            double min_a, max_a, min_b, max_b;
            project_polygon(a, nx, ny, min_a, max_a);
            project_polygon(b, nx, ny, min_b, max_b);
            if (interval_distance(min_a, max_a, min_b, max_b) > 0) intersecting = false;
            double vp = dot2(nx, ny, dax - dbx, day - dby);
            if (vp < 0)
                min_a += vp;
            else
                max_a += vp;
            double distance = interval_distance(min_a, max_a, min_b, max_b);
            if (distance > 0) will_intersect = false;
            if (!intersecting && !will_intersect) break;  // leaves the inner loop only
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
    trx = will_intersect ? min_distance * tax : 0.0;
    try_ = will_intersect ? min_distance * tay : 0.0;
}

// vehicle/controller.py:326-344 speed_to_index (np.round: half to even)
__device__ __forceinline__ int speed_to_index(const HwyHighwayParams& P, double speed) {
    int n = P.n_target_speeds;
    double x = (speed - P.target_speeds[0]) / (P.target_speeds[n - 1] - P.target_speeds[0]);
    return (int)clipd(rint(x * (n - 1)), 0.0, (double)(n - 1));
}

// ------------------------------------------------------------------ observation
// envs/common/observation.py:234-276 KinematicObservation.observe (presence,x,y,vx,vy; order
// "sorted") with road/road.py:421-450 close_objects_to and kinematics.py:237-261 to_dict.
// Requires sm.{x,y,c,s,v,lane[0]} published; `i` is the vehicle slot of the calling thread.
template <int TPE>
__device__ __forceinline__ void kinematics_observe(const HwyHighwayParams& P, EnvShared<TPE>& sm,
                                                   int i, float* __restrict__ obs_env) {
    const int V = P.n_vehicles, K = P.obs_vehicles_count;
    const HwyStraightLane& Le = P.lanes[sm.lane[0]];
    const double ex = sm.x[0], ey = sm.y[0];
    const double evx = sm.v[0] * sm.c[0], evy = sm.v[0] * sm.s[0];
    double key = INFINITY;
    if (i > 0 && i < V) {
        bool ok = norm2(sm.x[i] - ex, sm.y[i] - ey) < P.perception_distance;
        double d = lane_s(Le, sm.x[i], sm.y[i]) - lane_s(Le, ex, ey);
        ok = ok && (P.obs_see_behind || -2 * kVehLength < d);
        if (ok) key = fabs(d);
    }
    sm.key[i] = key;
    env_sync<TPE>();
    // stable rank among valid candidates (python sorted() on |lane_distance_to|)
    int rank = 0, n_valid = 0;
    for (int u = 1; u < V; ++u) {
        double ku = sm.key[u];
        n_valid += ku < INFINITY;
        rank += (ku < key) || (ku == key && u < i);
    }
    const double xr = 5.0 * kMaxSpeed, yr = 4.0 * P.lanes_count, vr = 2 * kMaxSpeed;
    int row = -1;
    double r1 = 0, r2 = 0, r3 = 0, r4 = 0;
    if (i == 0) {
        row = 0;
        r1 = ex;
        r2 = ey;
        r3 = evx;
        r4 = evy;
    } else if (key < INFINITY && rank < K - 1) {
        row = rank + 1;
        r1 = sm.x[i];
        r2 = sm.y[i];
        r3 = sm.v[i] * sm.c[i];
        r4 = sm.v[i] * sm.s[i];
        if (!P.obs_absolute) {
            r1 -= ex;
            r2 -= ey;
            r3 -= evx;
            r4 -= evy;
        }
    }
    if (row >= 0) {
        if (P.obs_normalize) {  // normalize_obs :207-232
            r1 = lmap(r1, -xr, xr, -1.0, 1.0);
            r2 = lmap(r2, -yr, yr, -1.0, 1.0);
            r3 = lmap(r3, -vr, vr, -1.0, 1.0);
            r4 = lmap(r4, -vr, vr, -1.0, 1.0);
            if (P.obs_clip) {
                r1 = clipd(r1, -1.0, 1.0);
                r2 = clipd(r2, -1.0, 1.0);
                r3 = clipd(r3, -1.0, 1.0);
                r4 = clipd(r4, -1.0, 1.0);
            }
        }
        float* o = obs_env + 5 * row;
        o[0] = 1.0f;
        o[1] = (float)r1;
        o[2] = (float)r2;
        o[3] = (float)r3;
        o[4] = (float)r4;
    }
    // zero padding of missing rows
    int filled = 1 + (n_valid < K - 1 ? n_valid : K - 1);
    if (i < K && i >= filled) {
        float* o = obs_env + 5 * i;
        o[0] = o[1] = o[2] = o[3] = o[4] = 0.0f;
    }
}

// ------------------------------------------------------------------ state I/O
struct VehicleRegs {
    double x, y, heading, speed, target_speed, timer, delta, imp_x, imp_y;
    int meta;
};

__device__ __forceinline__ void load_vehicle(const HwyHighwayState& S, size_t slot, VehicleRegs& r) {
    double2 a = reinterpret_cast<const double2*>(S.pos)[slot];
    double2 b = reinterpret_cast<const double2*>(S.hs)[slot];
    double2 c = reinterpret_cast<const double2*>(S.tt)[slot];
    double2 d = reinterpret_cast<const double2*>(S.imp)[slot];
    r.x = a.x;
    r.y = a.y;
    r.heading = b.x;
    r.speed = b.y;
    r.target_speed = c.x;
    r.timer = c.y;
    r.imp_x = d.x;
    r.imp_y = d.y;
    r.delta = S.delta[slot];
    r.meta = S.meta[slot];
}
__device__ __forceinline__ void store_vehicle(const HwyHighwayState& S, size_t slot,
                                              const VehicleRegs& r) {
    reinterpret_cast<double2*>(S.pos)[slot] = make_double2(r.x, r.y);
    reinterpret_cast<double2*>(S.hs)[slot] = make_double2(r.heading, r.speed);
    reinterpret_cast<double2*>(S.tt)[slot] = make_double2(r.target_speed, r.timer);
    reinterpret_cast<double2*>(S.imp)[slot] = make_double2(r.imp_x, r.imp_y);
    S.meta[slot] = r.meta;  // delta never changes during a step
}

__device__ __forceinline__ int meta_lane(int m) { return (m >> HWY_META_LANE_SHIFT) & 0xff; }
__device__ __forceinline__ int meta_target(int m) { return (m >> HWY_META_TARGET_SHIFT) & 0xff; }
__device__ __forceinline__ int meta_kind(int m) { return (m >> HWY_META_KIND_SHIFT) & 3; }
__device__ __forceinline__ int meta_set_lane(int m, int l) {
    return (m & ~(0xff << HWY_META_LANE_SHIFT)) | (l << HWY_META_LANE_SHIFT);
}
__device__ __forceinline__ int meta_set_target(int m, int l) {
    return (m & ~(0xff << HWY_META_TARGET_SHIFT)) | (l << HWY_META_TARGET_SHIFT);
}

template <int TPE>
__device__ __forceinline__ void publish(EnvShared<TPE>& sm, int i, const VehicleRegs& r) {
    double sn, cs;
    sincos(r.heading, &sn, &cs);
    sm.x[i] = r.x;
    sm.y[i] = r.y;
    sm.c[i] = cs;
    sm.s[i] = sn;
    sm.v[i] = r.speed;
    // getattr(ego_vehicle, "target_speed", 0): a plain Vehicle has none (behavior.py:172)
    sm.ts[i] = meta_kind(r.meta) == HWY_KIND_VEHICLE ? 0.0 : r.target_speed;
    sm.lane[i] = (unsigned char)meta_lane(r.meta);
    sm.tgt[i] = (unsigned char)meta_target(r.meta);
}

template <int TPE>
__device__ __forceinline__ void set_bit(u64 (&m)[(TPE + 63) / 64], int i) {
    atomicOr(&m[i >> 6], 1ull << (i & 63));
}

// ------------------------------------------------------------------ the step kernel
template <int TPE>
__global__ void __launch_bounds__(TPE == 32 ? 128 : TPE)
highway_step_kernel(const HwyHighwayParams P, const HwyHighwayState S,
                    const int32_t* __restrict__ action_i, const float* __restrict__ action_f,
                    float* __restrict__ obs, double* __restrict__ reward,
                    uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated,
                    double* __restrict__ info_speed, uint8_t* __restrict__ info_crashed) {
    constexpr int EPB = TPE == 32 ? 4 : 1;  // envs per block (a warp per env when TPE == 32)
    constexpr int NW = (TPE + 63) / 64;
    __shared__ EnvShared<TPE> smem[EPB];
    const int sub = threadIdx.x / TPE;
    const int i = threadIdx.x % TPE;
    const int env = blockIdx.x * EPB + sub;
    // Whole blocks (TPE >= 64: EPB envs share __syncthreads) must stay together: out-of-range
    // envs clamp to the last env and skip their stores.
    const bool env_ok = env < S.n_envs;
    const int e = env_ok ? env : S.n_envs - 1;
    EnvShared<TPE>& sm = smem[sub];
    const int V = P.n_vehicles;
    const bool active = i < V;
    const size_t slot = (size_t)e * S.vp + (active ? i : 0);

    VehicleRegs r;
    load_vehicle(S, slot, r);
    const int kind = meta_kind(r.meta);
    int speed_index = (i == 0) ? S.speed_index[e] : 0;
    double act_steer = 0.0, act_accel = 0.0;

    if (active) {
        sm.flags[i] = (unsigned char)(((r.meta & HWY_META_CHECK_COLLISIONS) ? 1 : 0) |
                                      (kind != HWY_KIND_VEHICLE ? 2 : 0));
        publish(sm, i, r);
    }
    env_sync<TPE>();

    const int frames = P.simulation_frequency / P.policy_frequency;
    const double dt = 1.0 / P.simulation_frequency;
    const double diag = sqrt(kVehLength * kVehLength + kVehWidth * kVehWidth);

    for (int frame = 0; frame < frames; ++frame) {
        // ---- action_type.act(action) on the first frame (abstract.py:294-304)
        if (frame == 0) {
            if (i == 0) {
                if (kind == HWY_KIND_MDP) {
                    // MDPVehicle.act (controller.py:295-315) + ControlledVehicle.act lane part
                    // (:99-124); labels action.py:204.  follow_road (:135-143) cannot change the
                    // target on the single-road highway graph (next_lane hits KeyError,
                    // road/road.py:129-130).
                    int a = action_i[e];
                    if (a == 3 || a == 4) {
                        int idx = speed_to_index(P, r.speed) + (a == 3 ? 1 : -1);
                        idx = max(0, min(idx, P.n_target_speeds - 1));
                        speed_index = idx;
                        r.target_speed = P.target_speeds[idx];
                        sm.ts[0] = r.target_speed;
                    } else if (a == 0 || a == 2) {
                        int id = meta_target(r.meta) + (a == 2 ? 1 : -1);
                        id = max(0, min(id, P.lanes_count - 1));
                        if (lane_reachable(P.lanes[id], r.x, r.y)) {
                            r.meta = meta_set_target(r.meta, id);
                            sm.tgt[0] = (unsigned char)id;
                        }
                    }
                } else {
                    // ContinuousAction.get_action/act (action.py:136-162): Box is float32 and
                    // lmap (utils.py:31-33) stays in float32 (NEP 50 weak python scalars)
                    float a0 = action_f[2 * (size_t)e], a1 = action_f[2 * (size_t)e + 1];
                    if (P.act_clip) {
                        a0 = fminf(fmaxf(a0, -1.0f), 1.0f);
                        a1 = fminf(fmaxf(a1, -1.0f), 1.0f);
                    }
                    float acc = __fadd_rn((float)P.acc_lo,
                                          __fdiv_rn(__fmul_rn(__fsub_rn(a0, -1.0f),
                                                              (float)(P.acc_hi - P.acc_lo)), 2.0f));
                    float st = __fadd_rn((float)P.steer_lo,
                                         __fdiv_rn(__fmul_rn(__fsub_rn(a1, -1.0f),
                                                             (float)(P.steer_hi - P.steer_lo)), 2.0f));
                    act_accel = (double)acc;
                    act_steer = (double)st;
                }
            }
            env_sync<TPE>();
        }

        // ---- Road.act() (road/road.py:464-467), phase A: own-lane IDM + lane-change policy
        const int lane = meta_lane(r.meta);
        const int tgt0 = meta_target(r.meta);
        const bool crashed = (r.meta & HWY_META_CRASHED) != 0;
        const bool idm_active = active && kind == HWY_KIND_IDM && !crashed;  // behavior.py:102-103
        int tgt1 = tgt0;
        bool is_mid = false;
        double acc = 0.0;
        if (i < NW) {
            sm.mid[i] = 0;
            sm.changed[i] = 0;
            sm.aborted[i] = 0;
        }
        if (active) {
#pragma unroll
            for (int w = 0; w < NW; ++w) sm.geo[i][w] = 0;
        }
        env_sync<TPE>();
        if (idm_active) {
            int f_own, r_own;
            neighbours(sm, P.lanes[lane], V, i, f_own, r_own);
            acc = idm_acceleration(P, sm, r.delta, i, f_own);  // behavior.py:115-120
            if (lane != tgt0) {
                // change_lane_policy, ongoing change (behavior.py:229-244): geometric part of the
                // abort scan, 0 < d < d*; the target-lane conditions are resolved in order below.
                is_mid = true;
                const HwyStraightLane& L = P.lanes[lane];
                double s_i = lane_s(L, r.x, r.y);
                u64 g[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) g[w] = 0;
                for (int v = 0; v < V; ++v) {
                    if (v == i || !(sm.flags[v] & 2)) continue;
                    double d = lane_s(L, sm.x[v], sm.y[v]) - s_i;
                    double d_star = desired_gap(P, sm, i, v);
                    if (0 < d && d < d_star) g[v >> 6] |= 1ull << (v & 63);
                }
#pragma unroll
                for (int w = 0; w < NW; ++w) sm.geo[i][w] = g[w];
                set_bit<TPE>(sm.mid, i);
            } else if (P.lane_change_delay < r.timer) {  // utils.do_every (utils.py:27-28)
                r.timer = 0.0;
                // side_lanes (road/road.py:200-211): id-1 then id+1; no break => last wins
                for (int k = 0; k < 2; ++k) {
                    int cand = k == 0 ? lane - 1 : lane + 1;
                    if (cand < 0 || cand > P.lanes_count - 1) continue;
                    if (!lane_reachable(P.lanes[cand], r.x, r.y)) continue;
                    if (fabs(r.speed) < 1) continue;
                    if (mobil(P, sm, V, i, r.delta, cand, f_own, r_own)) tgt1 = cand;
                }
                if (tgt1 != tgt0) set_bit<TPE>(sm.changed, i);
            }
        }
        if (active) sm.tgt1[i] = (unsigned char)tgt1;
        env_sync<TPE>();

        // ---- ordered fix-up of the Gauss-Seidel abort scan (one thread per env).  Vehicles act
        // in list order; vehicle i sees the NEW target of every j < i and the OLD one of j > i.
        if (i == 0 && (sm.mid[0] | (NW > 1 ? sm.mid[NW - 1] : 0ull))) {
            for (int l = 0; l < P.lanes_count; ++l)
#pragma unroll
                for (int w = 0; w < NW; ++w) sm.tm[l][w] = sm.lane_ne[l][w] = 0;
            for (int v = 0; v < V; ++v) {
                u64 bit = 1ull << (v & 63);
                sm.tm[sm.tgt[v]][v >> 6] |= bit;
                for (int l = 0; l < P.lanes_count; ++l)
                    if (sm.lane[v] != l) sm.lane_ne[l][v >> 6] |= bit;
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                u64 ev = sm.mid[w] | sm.changed[w];
                while (ev) {
                    int b = __ffsll((long long)ev) - 1;
                    ev &= ev - 1;
                    int v = w * 64 + b;
                    u64 bit = 1ull << b;
                    int t0 = sm.tgt[v];
                    if (sm.changed[w] & bit) {  // MOBIL decision becomes visible to later vehicles
                        sm.tm[t0][w] &= ~bit;
                        sm.tm[sm.tgt1[v]][w] |= bit;
                    } else {
                        u64 hit = 0;
#pragma unroll
                        for (int w2 = 0; w2 < NW; ++w2)
                            hit |= sm.geo[v][w2] & sm.tm[t0][w2] & sm.lane_ne[t0][w2];
                        if (hit) {  // behavior.py:241-243: target := current lane
                            sm.aborted[w] |= bit;
                            sm.tm[t0][w] &= ~bit;
                            sm.tm[sm.lane[v]][w] |= bit;
                        }
                    }
                }
            }
        }
        env_sync<TPE>();

        // ---- Road.act() phase B: steering + target-lane IDM with the final target
        if (active) {
            int tgt = tgt1;
            if (is_mid && (sm.aborted[i >> 6] >> (i & 63)) & 1) tgt = lane;
            if (idm_active) {
                double steering = steering_control(P.lanes[tgt], r.x, r.y, r.heading, r.speed);
                if (lane != tgt) {  // behavior.py:121-131
                    int f_t, r_t;
                    neighbours(sm, P.lanes[tgt], V, i, f_t, r_t);
                    acc = fmin(acc, idm_acceleration(P, sm, r.delta, i, f_t));
                }
                act_steer = steering;
                act_accel = clipd(acc, -P.acc_max, P.acc_max);
            } else if (kind == HWY_KIND_MDP) {
                // ControlledVehicle.act(None) (controller.py:126-133); runs even when crashed
                act_steer = steering_control(P.lanes[tgt], r.x, r.y, r.heading, r.speed);
                act_accel = kKpA * (r.target_speed - r.speed);  // speed_control :189-198
            }
            r.meta = meta_set_target(r.meta, tgt);
        }

        // ---- Road.step(dt): Vehicle.step (kinematics.py:130-177; IDMVehicle.step behavior.py:139-148)
        env_sync<TPE>();  // everyone is done reading the pre-step staging
        if (active) {
            if (kind == HWY_KIND_IDM) r.timer += dt;
            if (crashed) {  // clip_actions :155-168
                act_steer = 0.0;
                act_accel = -1.0 * r.speed;
            }
            if (r.speed > kMaxSpeed)
                act_accel = fmin(act_accel, 1.0 * (kMaxSpeed - r.speed));
            else if (r.speed < kMinSpeed)
                act_accel = fmax(act_accel, 1.0 * (kMinSpeed - r.speed));
            double beta = atan(0.5 * tan(act_steer));
            double sn, cs;
            sincos(r.heading + beta, &sn, &cs);
            double vx = r.speed * cs, vy = r.speed * sn;
            r.x += vx * dt;
            r.y += vy * dt;
            if (r.meta & HWY_META_HAS_IMPACT) {
                r.x += r.imp_x;
                r.y += r.imp_y;
                r.meta = (r.meta | HWY_META_CRASHED) & ~HWY_META_HAS_IMPACT;
            }
            r.heading += r.speed * sin(beta) / (kVehLength / 2) * dt;
            r.speed += act_accel * dt;
            int nl = closest_lane(P, r.x, r.y, r.heading);  // on_state_update :170-177
            r.meta = meta_set_lane(r.meta, nl);
            if (kind == HWY_KIND_VEHICLE) r.meta = meta_set_target(r.meta, nl);  // schema: mirrors lane
            publish(sm, i, r);
        }
        env_sync<TPE>();

        // ---- Road.step collision sweep (road/road.py:477-481; objects.py:92-138).  Thread i
        // visits its partners in ascending order, so the surviving impact is the one written by
        // the largest partner index, exactly as the reference's (i < j) double loop leaves it.
        if (active) {
            const bool cc_i = sm.flags[i] & 1;
            for (int j = 0; j < V; ++j) {
                if (j == i || !(cc_i || (sm.flags[j] & 1))) continue;
                int a = i < j ? i : j, b = i < j ? j : i;
                double dist = norm2(sm.x[b] - sm.x[a], sm.y[b] - sm.y[a]);
                if (dist > (diag + diag) / 2 + sm.v[a] * dt) continue;
                double pa[5][2], pb[5][2];
                polygon(sm, a, pa);
                polygon(sm, b, pb);
                bool inter, will;
                double trx, try_;
                polygons_intersecting(pa, pb, sm.v[a] * sm.c[a] * dt, sm.v[a] * sm.s[a] * dt,
                                      sm.v[b] * sm.c[b] * dt, sm.v[b] * sm.s[b] * dt, inter, will,
                                      trx, try_);
                if (will) {
                    r.imp_x = i == a ? trx / 2 : -trx / 2;
                    r.imp_y = i == a ? try_ / 2 : -try_ / 2;
                    r.meta |= HWY_META_HAS_IMPACT;
                }
                if (inter) r.meta |= HWY_META_CRASHED;
            }
        }
        // no barrier needed here: the next frame's first shared writes (mid/geo/tgt1) touch
        // arrays the sweep does not read, and a barrier follows them.
    }

    // ---- epilogue: state back to HBM, observation, reward, termination
    if (active && env_ok) store_vehicle(S, slot, r);
    env_sync<TPE>();
    float* obs_env = obs + (size_t)e * P.obs_vehicles_count * 5;
    if (env_ok) kinematics_observe(P, sm, i, obs_env);
    else env_sync<TPE>();
    if (i == 0 && env_ok) {
        // envs/highway_env.py:100-151
        const int lane = meta_lane(r.meta);
        const HwyStraightLane& L = P.lanes[lane];
        int rl = kind == HWY_KIND_VEHICLE ? lane : meta_target(r.meta);
        double forward_speed = r.speed * sm.c[0];
        double scaled_speed = lmap(forward_speed, P.reward_speed_lo, P.reward_speed_hi, 0.0, 1.0);
        double es, elat;
        lane_local(L, r.x, r.y, es, elat);
        bool on_road = lane_on(L, es, elat, 0.0);
        bool is_crashed = (r.meta & HWY_META_CRASHED) != 0;
        int nl1 = P.lanes_count - 1 > 1 ? P.lanes_count - 1 : 1;
        double rew = 0.0;
        rew = rew + P.collision_reward * (is_crashed ? 1.0 : 0.0);
        rew = rew + P.right_lane_reward * ((double)rl / (double)nl1);
        rew = rew + P.high_speed_reward * clipd(scaled_speed, 0.0, 1.0);
        rew = rew + 0.0 * (on_road ? 1.0 : 0.0);
        if (P.normalize_reward)
            rew = lmap(rew, P.collision_reward, P.high_speed_reward + P.right_lane_reward, 0.0, 1.0);
        rew *= on_road ? 1.0 : 0.0;
        double t = S.time[e] + 1.0 / P.policy_frequency;  // abstract.py:274
        S.time[e] = t;
        S.speed_index[e] = speed_index;
        reward[e] = rew;
        terminated[e] = (uint8_t)(is_crashed || (P.offroad_terminal && !on_road));
        truncated[e] = (uint8_t)(t >= P.duration);
        if (info_speed) info_speed[e] = r.speed;  // abstract.py:200-217 _info
        if (info_crashed) info_crashed[e] = (uint8_t)is_crashed;
    }
}

// ------------------------------------------------------------------ observe-only kernel
template <int TPE>
__global__ void __launch_bounds__(TPE == 32 ? 128 : TPE)
highway_observe_kernel(const HwyHighwayParams P, const HwyHighwayState S,
                       const uint8_t* __restrict__ mask_a, const uint8_t* __restrict__ mask_b,
                       int use_mask, float* __restrict__ obs) {
    constexpr int EPB = TPE == 32 ? 4 : 1;
    __shared__ EnvShared<TPE> smem[EPB];
    const int sub = threadIdx.x / TPE, i = threadIdx.x % TPE;
    const int env = blockIdx.x * EPB + sub;
    const bool env_ok = env < S.n_envs;
    const int e = env_ok ? env : S.n_envs - 1;
    EnvShared<TPE>& sm = smem[sub];
    const bool active = i < P.n_vehicles;
    VehicleRegs r;
    load_vehicle(S, (size_t)e * S.vp + (active ? i : 0), r);
    if (active) publish(sm, i, r);
    env_sync<TPE>();
    const bool wanted = env_ok && (!use_mask || (mask_a && mask_a[e]) || (mask_b && mask_b[e]));
    if (wanted)
        kinematics_observe(P, sm, i, obs + (size_t)e * P.obs_vehicles_count * 5);
    else
        env_sync<TPE>();
}

// ------------------------------------------------------------------ reset kernel
// HighwayEnv._create_road/_create_vehicles (envs/highway_env.py:55-98,177-182) with
// Vehicle.create_random (vehicle/kinematics.py:50-104), IDMVehicle.__init__ timer and
// randomize_behavior (behavior.py:64-69), MDPVehicle.__init__ (controller.py:284-293).
// The spawn is a sequential chain on the env's PCG64 stream => one thread per env.
__global__ void __launch_bounds__(128)
highway_reset_kernel(const HwyHighwayParams P, const HwyHighwayState S,
                     const uint8_t* __restrict__ mask_a, const uint8_t* __restrict__ mask_b,
                     int use_mask) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n_envs) return;
    if (use_mask && !((mask_a && mask_a[e]) || (mask_b && mask_b[e]))) return;
    const int n = S.n_envs;
    Pcg64 g;
    g.s_hi = S.rng[0 * (size_t)n + e];
    g.s_lo = S.rng[1 * (size_t)n + e];
    g.i_hi = S.rng[2 * (size_t)n + e];
    g.i_lo = S.rng[3 * (size_t)n + e];
    u64 w4 = S.rng[4 * (size_t)n + e];
    g.has32 = (uint32_t)(w4 >> 32);
    g.u32 = (uint32_t)w4;

    double x_max = 0.0;  // running max of the longitudinal coordinates of spawned vehicles
    bool lanes_aligned = true;  // all lanes share origin-x and direction => s is lane independent
    for (int l = 1; l < P.lanes_count; ++l)
        lanes_aligned = lanes_aligned && P.lanes[l].start_x == P.lanes[0].start_x &&
                        P.lanes[l].dir_x == P.lanes[0].dir_x && P.lanes[l].dir_y == 0.0 &&
                        P.lanes[0].dir_y == 0.0;
    double2* pos = reinterpret_cast<double2*>(S.pos);
    double2* hs = reinterpret_cast<double2*>(S.hs);
    double2* tt = reinterpret_cast<double2*>(S.tt);
    double2* imp = reinterpret_cast<double2*>(S.imp);
    const size_t base = (size_t)e * S.vp;
    int ego_speed_index = -1;
    for (int v = 0; v < P.n_vehicles; ++v) {
        const bool is_ego = v == 0;
        // choice(list(graph.keys())) / choice(list(graph[_from].keys())): single element => no draw
        int id = (is_ego && P.initial_lane_id >= 0) ? P.initial_lane_id : g.choice(P.lanes_count);
        const HwyStraightLane& L = P.lanes[id];
        double speed = is_ego ? P.ego_speed : g.uniform(0.7 * L.speed_limit, 0.8 * L.speed_limit);
        double spacing = is_ego ? P.ego_spacing : 1 / P.vehicles_density;
        double default_spacing = 12 + 1.0 * speed;
        double offset = spacing * default_spacing * P.spawn_exp;
        double x0;
        if (v > 0) {
            if (lanes_aligned) {
                x0 = x_max;
            } else {  // np.max over lane.local_coordinates(v.position)[0] on the chosen lane
                x0 = lane_s(L, pos[base].x, pos[base].y);
                for (int j = 1; j < v; ++j) x0 = fmax(x0, lane_s(L, pos[base + j].x, pos[base + j].y));
            }
        } else {
            x0 = 3 * offset;
        }
        x0 += offset * g.uniform(0.9, 1.1);
        // lane.position(x0, 0), lane.heading_at(x0)  (road/lane.py:192-200)
        double px = (L.start_x + x0 * L.dir_x) + 0.0 * L.lat_x;
        double py = (L.start_y + x0 * L.dir_y) + 0.0 * L.lat_y;
        double heading = L.heading;
        double s_here = lane_s(L, px, py);
        x_max = v == 0 ? s_here : fmax(x_max, s_here);
        int lane = closest_lane(P, px, py, heading);  // RoadObject.__init__ objects.py:46-50
        double target_speed = speed;                   // `target_speed or self.speed`
        double timer = 0.0, delta = 4.0;
        int kind, cc;
        if (is_ego) {
            cc = 1;
            if (P.action_type == 0) {
                kind = HWY_KIND_MDP;
                ego_speed_index = speed_to_index(P, target_speed);
                target_speed = P.target_speeds[ego_speed_index];
            } else {
                kind = HWY_KIND_VEHICLE;
            }
        } else {
            kind = HWY_KIND_IDM;
            cc = P.others_check_collisions;
            timer = py_mod_pos((px + py) * kPi, P.lane_change_delay);  // behavior.py:64
            delta = g.uniform(P.delta_lo, P.delta_hi);                 // behavior.py:66-69
        }
        pos[base + v] = make_double2(px, py);
        hs[base + v] = make_double2(heading, speed);
        tt[base + v] = make_double2(target_speed, timer);
        imp[base + v] = make_double2(0.0, 0.0);
        S.delta[base + v] = delta;
        S.meta[base + v] = (lane << HWY_META_LANE_SHIFT) | (lane << HWY_META_TARGET_SHIFT) |
                           (cc ? HWY_META_CHECK_COLLISIONS : 0) | (kind << HWY_META_KIND_SHIFT) |
                           HWY_META_PRESENT;
    }
    S.speed_index[e] = ego_speed_index;
    S.time[e] = 0.0;
    S.rng[0 * (size_t)n + e] = g.s_hi;
    S.rng[1 * (size_t)n + e] = g.s_lo;
    S.rng[2 * (size_t)n + e] = g.i_hi;
    S.rng[3 * (size_t)n + e] = g.i_lo;
    S.rng[4 * (size_t)n + e] = ((u64)g.has32 << 32) | g.u32;
}

}  // namespace hwy

// ====================================================================== C ABI
namespace {
thread_local char g_err[512] = "";
thread_local unsigned long long g_launches = 0;

int fail(const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return 1;
}
int check_launch(const char* what) {
    ++g_launches;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(err));
        return 1;
    }
    return 0;
}
int validate(const HwyHighwayParams* p, const HwyHighwayState* s) {
    if (!p || !s) return fail("%s", "null params/state");
    if (p->n_vehicles < 1 || p->n_vehicles > HWY_MAX_VEHICLES) return fail("%s", "n_vehicles out of range");
    if (p->lanes_count < 1 || p->lanes_count > HWY_MAX_LANES) return fail("%s", "lanes_count out of range");
    if (p->obs_vehicles_count < 1 || p->obs_vehicles_count > HWY_MAX_OBS_VEHICLES)
        return fail("%s", "obs_vehicles_count out of range");
    if (p->n_target_speeds < 1 || p->n_target_speeds > HWY_MAX_TARGET_SPEEDS)
        return fail("%s", "n_target_speeds out of range");
    if (p->simulation_frequency < 1 || p->policy_frequency < 1 ||
        p->simulation_frequency < p->policy_frequency)
        return fail("%s", "bad simulation/policy frequency");
    if (s->n_envs < 1) return fail("%s", "n_envs < 1");
    if (s->vp < p->n_vehicles || (s->vp & 1)) return fail("%s", "slot stride must be even and >= n_vehicles");
    if (!s->pos || !s->hs || !s->tt || !s->imp || !s->delta || !s->meta || !s->speed_index ||
        !s->time || !s->rng)
        return fail("%s", "null state pointer");
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count < 1) {
        cudaGetLastError();
        return fail("%s", "no CUDA device: this library has no CPU fallback");
    }
    return 0;
}
int tpe_for(int n_vehicles) { return n_vehicles <= 32 ? 32 : (n_vehicles <= 64 ? 64 : 128); }

struct Grid {
    int blocks, threads;
};
Grid grid_for(int tpe, int n_envs) {
    int epb = tpe == 32 ? 4 : 1;
    return Grid{(n_envs + epb - 1) / epb, tpe * epb};
}

int launch_observe(const HwyHighwayParams* p, const HwyHighwayState* s, const uint8_t* mask_a,
                   const uint8_t* mask_b, int use_mask, float* obs, cudaStream_t st) {
    int tpe = tpe_for(p->n_vehicles);
    Grid g = grid_for(tpe, s->n_envs);
    if (tpe == 32)
        hwy::highway_observe_kernel<32><<<g.blocks, g.threads, 0, st>>>(*p, *s, mask_a, mask_b, use_mask, obs);
    else if (tpe == 64)
        hwy::highway_observe_kernel<64><<<g.blocks, g.threads, 0, st>>>(*p, *s, mask_a, mask_b, use_mask, obs);
    else
        hwy::highway_observe_kernel<128><<<g.blocks, g.threads, 0, st>>>(*p, *s, mask_a, mask_b, use_mask, obs);
    return check_launch("highway_observe_kernel");
}

int launch_reset(const HwyHighwayParams* p, const HwyHighwayState* s, const uint8_t* mask_a,
                 const uint8_t* mask_b, int use_mask, cudaStream_t st) {
    int blocks = (s->n_envs + 127) / 128;
    hwy::highway_reset_kernel<<<blocks, 128, 0, st>>>(*p, *s, mask_a, mask_b, use_mask);
    return check_launch("highway_reset_kernel");
}
}  // namespace

extern "C" {

int hwy_abi_version(void) { return HWY_ABI_VERSION; }
const char* hwy_last_error(void) { return g_err; }
uint64_t hwy_launch_count(void) { return g_launches; }
int hwy_highway_slot_stride(int n_vehicles) { return (n_vehicles + 1) & ~1; }

int hwy_highway_observe(const HwyHighwayParams* p, const HwyHighwayState* s, float* obs, void* stream) {
    if (validate(p, s)) return 1;
    if (!obs) return fail("%s", "obs is null");
    return launch_observe(p, s, nullptr, nullptr, 0, obs, (cudaStream_t)stream);
}

int hwy_highway_reset(const HwyHighwayParams* p, const HwyHighwayState* s, const uint8_t* mask,
                      float* obs, void* stream) {
    if (validate(p, s)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    int use_mask = mask != nullptr;
    if (launch_reset(p, s, mask, nullptr, use_mask, st)) return 1;
    if (obs) return launch_observe(p, s, mask, nullptr, use_mask, obs, st);
    return 0;
}

int hwy_highway_autoreset(const HwyHighwayParams* p, const HwyHighwayState* s,
                          const uint8_t* terminated, const uint8_t* truncated, float* obs,
                          void* stream) {
    if (validate(p, s)) return 1;
    if (!terminated || !truncated || !obs) return fail("%s", "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (launch_reset(p, s, terminated, truncated, 1, st)) return 1;
    return launch_observe(p, s, terminated, truncated, 1, obs, st);
}

int hwy_highway_step(const HwyHighwayParams* p, const HwyHighwayState* s, const int32_t* action_i,
                     const float* action_f, float* obs, double* reward, uint8_t* terminated,
                     uint8_t* truncated, double* info_speed, uint8_t* info_crashed, int autoreset,
                     float* final_obs, void* stream) {
    if (validate(p, s)) return 1;
    if (!obs || !reward || !terminated || !truncated) return fail("%s", "null output pointer");
    if (p->action_type == 0 && !action_i) return fail("%s", "DiscreteMetaAction needs action_i");
    if (p->action_type == 1 && !action_f) return fail("%s", "ContinuousAction needs action_f");
    if (autoreset != HWY_AUTORESET_DISABLED && autoreset != HWY_AUTORESET_SAME_STEP)
        return fail("%s", "unknown autoreset mode");
    cudaStream_t st = (cudaStream_t)stream;
    int tpe = tpe_for(p->n_vehicles);
    Grid g = grid_for(tpe, s->n_envs);
    if (tpe == 32)
        hwy::highway_step_kernel<32><<<g.blocks, g.threads, 0, st>>>(*p, *s, action_i, action_f, obs,
                                                                    reward, terminated, truncated, info_speed,
            info_crashed);
    else if (tpe == 64)
        hwy::highway_step_kernel<64><<<g.blocks, g.threads, 0, st>>>(*p, *s, action_i, action_f, obs,
                                                                    reward, terminated, truncated, info_speed,
            info_crashed);
    else
        hwy::highway_step_kernel<128><<<g.blocks, g.threads, 0, st>>>(*p, *s, action_i, action_f, obs,
                                                                     reward, terminated, truncated, info_speed,
            info_crashed);
    if (check_launch("highway_step_kernel")) return 1;
    if (autoreset == HWY_AUTORESET_SAME_STEP) {
        if (final_obs) {
            size_t bytes = (size_t)s->n_envs * p->obs_vehicles_count * 5 * sizeof(float);
            cudaError_t err = cudaMemcpyAsync(final_obs, obs, bytes, cudaMemcpyDeviceToDevice, st);
            if (err != cudaSuccess) return fail("final_obs copy: %s", cudaGetErrorString(err));
        }
        // envs with terminated | truncated restart from their own RNG stream; obs := reset obs
        if (launch_reset(p, s, terminated, truncated, 1, st)) return 1;
        return launch_observe(p, s, terminated, truncated, 1, obs, st);
    }
    return 0;
}

}  // extern "C"
