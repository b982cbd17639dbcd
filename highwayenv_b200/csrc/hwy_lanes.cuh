// hwy_lanes.cuh — lane geometry of a general road network (road/lane.py: StraightLane, SineLane, CircularLane) and the
// shared-memory copy of the lane table, shared by the step kernels (hwy_network.cu) and the observation plugins
// (hwy_observe.cu).  Reference paths are relative to /root/reference/highway_env.
#pragma once
#include <cuda_runtime.h>

#include "../../include/hwyb200.h"
#include "hwy_device.cuh"
#include "hwy_math.cuh"

namespace hwynet {
using namespace hwy;

struct GraphShared {
    int n_lanes, n_nodes;
    HwyNetLane lanes[HWY_NET_MAX_LANES];
    int succ_count[HWY_NET_MAX_NODES];
    int succ[HWY_NET_MAX_NODES][HWY_NET_MAX_SUCC];
};

// ------------------------------------------------------------------ lanes (road/lane.py)
// local_coordinates: StraightLane :205-209, SineLane :285-289, CircularLane :351-358
static __device__ __noinline__ void lane_local(const HwyNetLane& L, double x, double y, double& s, double& lat) {
    if (L.type == HWY_LANE_CIRCULAR) {
        double ddx = x - L.cx, ddy = y - L.cy;
        double r = norm2(ddx, ddy);
        lat = L.direction * (L.radius - r);
        double phi = atan2(ddy, ddx);
        phi = L.start_phase + wrap_to_pi(phi - L.start_phase);
        s = L.direction * (phi - L.start_phase) * L.radius;
        return;
    }
    double ddx = x - L.sx, ddy = y - L.sy;
    double la = dot2(ddx, ddy, L.lx, L.ly);
    double lon = dot2(ddx, ddy, L.dx, L.dy);
    if (L.type == HWY_LANE_SINE) la = la - L.amplitude * m_sin(L.pulsation * lon + L.phase);
    s = lon;
    lat = la;
}
// on_lane(position, margin = 1) pre-test for the neighbour search (road/road.py:483-547 -> lane.py:100-118): false only
// when local_coordinates followed by on_lane CANNOT hold, decided without sqrt / atan2 / sin, so that the expensive
// exact evaluation runs for the few real candidates only (and, gathered in a pending mask, converged across the warp).
//   circular: |radius - r| <= gate compared on squares with a 1e-12 relative slack; and for arcs shorter than pi, a point
//     outside the arc's angular sector (cross products against the end points' radial vectors, margin towards "inside")
//     whose distance to the nearer end point exceeds gate + VEHICLE_LENGTH cannot have -LENGTH <= s < length + LENGTH:
//     at an angular excess of at most LENGTH / radius beyond an end, |p - end| <= |r - radius| + chord <= gate + LENGTH.
//     (`L` is the kernel's shared copy: stage_graph keeps the arc's end points in sx, sy, ex, ey.)
//   sine: |lateral| >= |straight lateral| - |amplitude| (|A sin| <= |A| and rounding is monotone);
//   straight / sine: the longitudinal test itself (s is the straight longitudinal on both).
__device__ __forceinline__ bool lane_maybe_on(const HwyNetLane& L, double x, double y) {
    const double gate = L.width / 2 + 1.0;
    if (L.type == HWY_LANE_CIRCULAR) {
        const double wx = x - L.cx, wy = y - L.cy;
        const double r2 = dot2(wx, wy, wx, wy);
        const double hi = L.radius + gate, lo = L.radius - gate;
        if (r2 > hi * hi * (1.0 + 1e-12)) return false;
        if (lo > 0.0 && r2 < lo * lo * (1.0 - 1e-12)) return false;
        if (fabs(L.end_phase - L.start_phase) < 3.0) {
            const double usx = L.sx - L.cx, usy = L.sy - L.cy, uex = L.ex - L.cx, uey = L.ey - L.cy;
            const double c_s = L.direction * (usx * wy - usy * wx);  // > 0: past the start point in travel direction
            const double c_e = L.direction * (wx * uey - wy * uex);  // > 0: before the end point
            if (c_s < -1e-6 || c_e < -1e-6) {
                const double ax = x - L.sx, ay = y - L.sy, bx = x - L.ex, by = y - L.ey;
                const double reach = gate + kLaneVehLength + 1e-6;
                if (fmin(dot2(ax, ay, ax, ay), dot2(bx, by, bx, by)) > reach * reach) return false;
            }
        }
        return true;
    }
    const double ddx = x - L.sx, ddy = y - L.sy;
    const double la = fabs(dot2(ddx, ddy, L.lx, L.ly));
    if (L.type == HWY_LANE_SINE ? la - fabs(L.amplitude) > gate : !(la <= gate)) return false;
    const double lon = dot2(ddx, ddy, L.dx, L.dy);
    return -kLaneVehLength <= lon && lon < L.length + kLaneVehLength;
}
// position: StraightLane :192-197, SineLane :268-273, CircularLane :338-342
static __device__ __noinline__ void lane_position(const HwyNetLane& L, double s, double lat, double& x, double& y) {
    if (L.type == HWY_LANE_CIRCULAR) {
        double phi = L.direction * s / L.radius + L.start_phase;
        double rr = L.radius - lat * L.direction;
        double sn, cs;
        m_sincos(phi, &sn, &cs);
        x = L.cx + rr * cs;
        y = L.cy + rr * sn;
        return;
    }
    if (L.type == HWY_LANE_SINE) lat = lat + L.amplitude * m_sin(L.pulsation * s + L.phase);
    x = (L.sx + s * L.dx) + lat * L.lx;
    y = (L.sy + s * L.dy) + lat * L.ly;
}
// heading_at: StraightLane :199-200, SineLane :275-280, CircularLane :344-347
static __device__ __noinline__ double lane_heading_at(const HwyNetLane& L, double s) {
    if (L.type == HWY_LANE_CIRCULAR) {
        double phi = L.direction * s / L.radius + L.start_phase;
        return phi + kPi / 2 * L.direction;
    }
    if (L.type == HWY_LANE_SINE) {
        double sn, cs;
        m_sincos(L.pulsation * s + L.phase, &sn, &cs);
        return L.heading + m_atan(L.amplitude * L.pulsation * cs);
    }
    return L.heading;
}
__device__ __forceinline__ double lane_s_of(const HwyNetLane& L, double x, double y) {
    double s, lat;
    lane_local(L, x, y, s, lat);
    return s;
}
__device__ __forceinline__ bool lane_on(const HwyNetLane& L, double s, double lat, double margin) {
    return fabs(lat) <= L.width / 2 + margin && -kLaneVehLength <= s && s < L.length + kLaneVehLength;
}
__device__ __forceinline__ bool lane_reachable(const HwyNetLane& L, double x, double y) {
    if (L.forbidden) return false;
    double s, lat;
    lane_local(L, x, y, s, lat);
    return fabs(lat) <= 2 * L.width && 0 <= s && s < L.length + kLaneVehLength;
}
// :127-130 distance
__device__ __forceinline__ double lane_distance(const HwyNetLane& L, double x, double y) {
    double s, r;
    lane_local(L, x, y, s, r);
    return fabs(r) + fmax(s - L.length, 0.0) + fmax(0.0 - s, 0.0);
}
// :132-147 distance_with_heading
__device__ __forceinline__ double lane_distance_with_heading(const HwyNetLane& L, double x, double y, double h,
                                                             double& s, double& r) {
    lane_local(L, x, y, s, r);
    double angle = fabs(wrap_to_pi(h - lane_heading_at(L, s)));
    return fabs(r) + fmax(s - L.length, 0.0) + fmax(0.0 - s, 0.0) + 1.0 * angle;
}

#define RT_FROM(e) ((e)&0xff)
#define RT_TO(e) (((e) >> 8) & 0xff)
#define RT_ID(e) ((((e) >> 16) & 0xff) - 1)

__device__ __forceinline__ int road_first(const GraphShared& g, int from, int to) {
    for (int k = 0; k < g.succ_count[from]; ++k) {
        int f = g.succ[from][k];
        if (g.lanes[f].to_node == to) return f;
    }
    return -1;
}

__device__ __forceinline__ double lane_distance_with_heading(const HwyNetLane& L, double x, double y,
                                                             double h) {
    double s, r;
    return lane_distance_with_heading(L, x, y, h, s, r);
}

__device__ __forceinline__ void stage_graph(GraphShared& gs, const HwyNetGraph* __restrict__ graph) {
    static_assert(sizeof(GraphShared) == sizeof(HwyNetGraph), "layout");
    const int* src = reinterpret_cast<const int*>(graph);
    int* dst = reinterpret_cast<int*>(&gs);
    for (int k = threadIdx.x; k < (int)(sizeof(HwyNetGraph) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
    __syncthreads();
    // CircularLane rows do not use the StraightLane fields: the shared copy keeps the arc's two end points there
    // (position(0, 0) and position(length, 0)), for the closest-lane pruning and the neighbour pre-test (closest_lane_prunable, lane_maybe_on)
    for (int l = threadIdx.x; l < gs.n_lanes; l += blockDim.x) {
        HwyNetLane& L = gs.lanes[l];
        if (L.type != HWY_LANE_CIRCULAR) continue;
        double sn, cs;
        sincos(L.start_phase, &sn, &cs);
        L.sx = L.cx + L.radius * cs;
        L.sy = L.cy + L.radius * sn;
        sincos(L.end_phase, &sn, &cs);
        L.ex = L.cx + L.radius * cs;
        L.ey = L.cy + L.radius * sn;
    }
    __syncthreads();
}

// get_closest_lane_index pruning (road/road.py:55-71): true when lane.distance_with_heading(position, heading)
// (road/lane.py:132-143) of L certainly exceeds `bd` (the distance to the lane the vehicle was on), decided without
// sqrt / atan2 / sin:
//   d = |lateral| + max(s - length, 0) + max(-s, 0) + |angle|   (all terms >= 0; fp addition of non-negative terms is
//   monotone, so dropping terms or replacing one by something smaller can only lower the sum).
// Straight: the first three terms themselves (exact).  Sine: |lateral| >= |straight lateral| - |amplitude| (1e-9 m
// covers the rounding of that subtraction), same longitudinal terms.  Circular: |lateral| = |radius - r| > bd, compared
// on squares (r^2 against (radius +- bd)^2 with a 1e-12 relative slack, far above the rounding of either side); and when
// the point lies outside the arc's angular sector (arcs shorter than pi; decided with cross products against the end
// points' radial vectors, with a margin that can only misjudge towards "inside"), |lateral| + the arc length beyond the
// violated end >= |p - q| + chord(q, end) >= |p - end| for its radial projection q, so the distance to the NEARER end
// point bounds d from below whichever end the reference's phase wrap picks (squares again, 1e-9 m + 1e-12 slack).
// A lane that is not pruned is evaluated exactly, so slack only costs time.
// `cache`: (cx, cy, r^2) of the last circular lane evaluated — the ring arcs of a roundabout share one centre.
struct CircleCache {
    double cx, cy, r2;
    bool valid;
};
__device__ __forceinline__ bool closest_lane_prunable(const HwyNetLane& L, double x, double y, double bd,
                                                      CircleCache& cache) {
    if (L.type == HWY_LANE_CIRCULAR) {
        const double wx = x - L.cx, wy = y - L.cy;
        if (!(cache.valid && cache.cx == L.cx && cache.cy == L.cy)) {
            cache.cx = L.cx;
            cache.cy = L.cy;
            cache.r2 = dot2(wx, wy, wx, wy);
            cache.valid = true;
        }
        const double hi = L.radius + bd, lo = L.radius - bd;
        if (cache.r2 > hi * hi * (1.0 + 1e-12)) return true;
        if (lo > 0.0 && cache.r2 < lo * lo * (1.0 - 1e-12)) return true;
        if (fabs(L.end_phase - L.start_phase) < 3.0) {
            const double usx = L.sx - L.cx, usy = L.sy - L.cy, uex = L.ex - L.cx, uey = L.ey - L.cy;
            const double c_s = L.direction * (usx * wy - usy * wx);  // > 0: past the start point in travel direction
            const double c_e = L.direction * (wx * uey - wy * uex);  // > 0: before the end point
            if (c_s < -1e-6 || c_e < -1e-6) {
                const double ax = x - L.sx, ay = y - L.sy, bx = x - L.ex, by = y - L.ey;
                const double reach = bd + 1e-9;
                if (fmin(dot2(ax, ay, ax, ay), dot2(bx, by, bx, by)) > reach * reach * (1.0 + 1e-12)) return true;
            }
        }
        return false;
    }
    const double ddx = x - L.sx, ddy = y - L.sy;
    double lat = fabs(dot2(ddx, ddy, L.lx, L.ly));
    if (L.type == HWY_LANE_SINE) lat = lat - fabs(L.amplitude) - 1e-9;
    const double lon = dot2(ddx, ddy, L.dx, L.dy);
    return lat + fmax(lon - L.length, 0.0) + fmax(0.0 - lon, 0.0) > bd;
}

}  // namespace hwynet
