/*
 * net_oracle.h — CPU restatement of the HighwayEnv hot path on a GENERAL road network
 * (StraightLane / SineLane / CircularLane, planned routes): roundabout-v0.
 * TEST INFRASTRUCTURE ONLY, same rules as hwy_oracle.h: never included, linked or called by
 * the product.  Scalar, sequential, per-vehicle; each function cites the reference
 * file:line it follows (paths relative to /root/reference/highway_env).  Pinned by
 * tests/test_net_oracle_golden.py against golden rollouts of the unmodified reference.
 */
#ifndef NET_ORACLE_H
#define NET_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NET_MAX_LANES 64
#define NET_MAX_NODES 64
#define NET_MAX_SUCC 6
#define NET_MAX_ROUTE 16
#define NET_MAX_TARGET_SPEEDS 8

#define NET_LANE_STRAIGHT 0 /* road/lane.py:159 */
#define NET_LANE_SINE 1     /* road/lane.py:236 */
#define NET_LANE_CIRCULAR 2 /* road/lane.py:311 */

#define NET_KIND_IDM 0
#define NET_KIND_MDP 1
#define NET_KIND_VEHICLE 2  /* vehicle/kinematics.py:13 Vehicle (ContinuousAction ego); with cfg->dynamical: BicycleVehicle */
#define NET_KIND_OBSTACLE 3 /* vehicle/objects.py:213-220 Obstacle: a static 2 x 2 m road object (road.objects);
                             * objects occupy the slots AFTER the vehicles */

#define NET_OBS_KINEMATICS 0 /* envs/common/observation.py:155 */
#define NET_OBS_OCCUPANCY 1  /* envs/common/observation.py:279 */
#define NET_OBS_TTC 2        /* envs/common/observation.py:115 */

/* One lane of RoadNetwork.graph[from][to][lane_id]; table order = graph enumeration order
 * (road/road.py:65-71). */
typedef struct NetLane {
    int32_t type, from_node, to_node, lane_id;
    int32_t road_first, road_count; /* table index of lane 0 of this road, lanes on the road */
    int32_t forbidden, priority;
    int32_t exit_lane, _pad; /* intersection: "il" in from-node name and "o" in to-node name */
    double width, speed_limit, length;
    double sx, sy, ex, ey, dx, dy, lx, ly, heading; /* StraightLane fields (lane.py:183-194) */
    double amplitude, pulsation, phase;             /* SineLane */
    double cx, cy, radius, start_phase, end_phase, direction; /* CircularLane */
} NetLane;

typedef struct NetGraph {
    int32_t n_lanes, n_nodes;
    NetLane lanes[NET_MAX_LANES];
    /* graph[node].keys() in insertion order: first-lane table index of each outgoing road */
    int32_t succ_count[NET_MAX_NODES];
    int32_t succ[NET_MAX_NODES][NET_MAX_SUCC];
} NetGraph;

typedef struct NetCfg {
    int32_t n_vehicles;
    int32_t simulation_frequency, policy_frequency;
    int32_t n_target_speeds;
    int32_t obs_type;
    int32_t obs_vehicles_count, obs_see_behind, obs_absolute, obs_normalize, obs_clip;
    int32_t ttc_horizon;
    int32_t normalize_reward;
    double duration;
    double target_speeds[NET_MAX_TARGET_SPEEDS];
    double obs_x_lo, obs_x_hi, obs_y_lo, obs_y_hi, obs_vx_lo, obs_vx_hi, obs_vy_lo, obs_vy_hi;
    double collision_reward, high_speed_reward, lane_change_reward; /* roundabout_env.py:30-34 */
    double acc_max, comfort_acc_max, comfort_acc_min, distance_wanted, time_wanted;
    double politeness, lane_change_min_acc_gain, lane_change_max_braking_imposed, lane_change_delay;
    double perception_distance;
    /* intersection-v0 (envs/intersection_env.py) */
    int32_t regulated;        /* RegulatedRoad (road/regulation.py) */
    int32_t action_mode;      /* 0: LANE_LEFT/IDLE/LANE_RIGHT/FASTER/SLOWER, 1: SLOWER/IDLE/FASTER (action.py:206) */
    int32_t reward_type;      /* 0 roundabout, 1 intersection */
    int32_t obs_features;     /* Kinematics: 5, or 7 with cos_h, sin_h */
    int32_t offroad_terminal;
    int32_t connected_lanes;  /* config["neighbour_vehicles_connected_lanes"] (abstract.py:26-37, road.py:509-529) */
    double arrived_reward, reward_speed_lo, reward_speed_hi;
    /* merge-v0 (envs/merge_env.py): reward_type 2 */
    double right_lane_reward, merging_speed_reward;
    int32_t merge_lane; /* table index of ("b", "c", 2), the lane whose slow vehicles are penalised */
    int32_t _pad3;
    /* two-way-v0 (envs/two_way_env.py): reward_type 3; u-turn-v0 (envs/u_turn_env.py): reward_type 4 */
    double left_lane_reward;
    double ego_pursuit_tau; /* u_turn_env.py:193 ego.PURSUIT_TAU = TAU_HEADING; 0: the class default */
    /* exit-v0 (envs/exit_env.py): reward_type 5 */
    double goal_reward;
    int32_t exit_lane_a, exit_lane_b; /* table indices of ("1","2",lanes_count) and ("2","exit",0): _is_success (:178-190) */
    /* ContinuousAction / DiscreteAction (envs/common/action.py:73-196): the controlled vehicle is a plain Vehicle, or
     * with `dynamical` a BicycleVehicle (vehicle/dynamics.py:33-160) */
    int32_t action_type;              /* 0 DiscreteMetaAction, 1 ContinuousAction (float32 [throttle, steering]) */
    int32_t act_clip, dynamical, obs_n_feat; /* obs_n_feat > 0: Kinematics columns from obs_feat[] (NET_FEAT_*) */
    double acc_lo, acc_hi, steer_lo, steer_hi;
    int32_t obs_feat[16], obs_feat_ranged[16];
    double obs_feat_lo[16], obs_feat_hi[16];
    int32_t obs_exit_lane;            /* ExitObservation (observation.py:624-675): table index of ("1","2",-1) whose
                                       * longitudinal coordinate replaces the ego row's x; -1: plain Kinematics */
    int32_t _pad4;
} NetCfg;

/* route entry: from | to << 8 | (lane_id + 1) << 16   (lane_id + 1 == 0: None) */
#define NET_ROUTE(from, to, id) ((from) | ((to) << 8) | (((id) + 1) << 16))

typedef struct NetState {
    double *x, *y, *heading, *speed, *target_speed, *timer, *delta, *impact_x, *impact_y;
    int32_t *lane, *target_lane, *kind, *crashed, *has_impact, *check_collisions;
    int32_t *route;     /* [V][NET_MAX_ROUTE] */
    int32_t *route_len; /* [V] */
    int32_t *speed_index; /* [1] */
    double *time;         /* [1] */
    int32_t *count;       /* [1] current number of vehicles (dynamic population); NULL: cfg->n_vehicles */
    int32_t *is_yielding; /* [V] RegulatedRoad yield flag */
    int32_t *road_steps;  /* [1] RegulatedRoad.steps */
    int32_t *no_lane_change; /* [V] IDMVehicle(enable_lane_change=False) (behavior.py:48-62,104-105); NULL: all enabled */
    double *lat_speed, *yaw_rate; /* [V] BicycleVehicle.lateral_speed / yaw_rate (dynamics.py:52-53); NULL: kinematic */
} NetState;

/* One AbstractEnv.step of a roundabout-v0 style env (MDPVehicle ego in slot 0). */
void net_step(const NetGraph *g, const NetCfg *c, NetState *s, int action, float *obs,
              double *reward, int32_t *terminated, int32_t *truncated);
void net_observe(const NetGraph *g, const NetCfg *c, const NetState *s, float *obs);
/* the same step with a ContinuousAction ego: action = float32 [throttle, steering] in [-1, 1] */
void net_step_continuous(const NetGraph *g, const NetCfg *c, NetState *s, const float *action, float *obs,
                         double *reward, int32_t *terminated, int32_t *truncated);
/* several controlled vehicles (MultiAgentAction / MultiAgentObservation, intersection-multi-agent-v0):
 * actions [n_agents], obs [n_agents][net_obs_size], speed_index [n_agents] in the state */
void net_step_agents(const NetGraph *g, const NetCfg *c, NetState *s, const int32_t *actions, int n_agents,
                     float *obs, double *reward, int32_t *terminated, int32_t *truncated, double *agents_reward,
                     int32_t *agents_terminated);
void net_observe_agents(const NetGraph *g, const NetCfg *c, const NetState *s, int n_agents, float *obs);
/* Road.act() + Road.step(dt) `substeps` times without an ego action (intersection warm-up) */
void net_substeps(const NetGraph *g, const NetCfg *c, NetState *s, int substeps);
/* has_arrived(vehicle) of envs/intersection_env.py:368-373 */
int net_has_arrived(const NetGraph *g, const NetState *s, int v);
int net_obs_size(const NetCfg *c);

/* entries for the reference's own known-answer tests */
void net_neighbours(const NetGraph *g, const NetCfg *c, const NetState *s, int veh, int lane_idx, int32_t *front,
                    int32_t *rear);
int net_rotated_rectangles_intersect(double c1x, double c1y, double l1, double w1, double a1, double c2x, double c2y,
                                     double l2, double w2, double a2);

/* ------------------------------------------------------------------ observation plugins on ANY road family
 * (envs/common/observation.py:772-794 observation_factory: every observation type works on every env).  The state
 * is the per-env NetState with V vehicles/objects; `ego` is the observer's slot. */
#define NET_FEAT_PRESENCE 0
#define NET_FEAT_X 1
#define NET_FEAT_Y 2
#define NET_FEAT_VX 3
#define NET_FEAT_VY 4
#define NET_FEAT_HEADING 5
#define NET_FEAT_COS_H 6
#define NET_FEAT_SIN_H 7
#define NET_FEAT_COS_D 8
#define NET_FEAT_SIN_D 9
#define NET_FEAT_LONG_OFF 10
#define NET_FEAT_LAT_OFF 11
#define NET_FEAT_ANG_OFF 12
#define NET_FEAT_ON_ROAD 13  /* OccupancyGrid only */
#define NET_FEAT_UNKNOWN 14  /* a feature name no Vehicle.to_dict key matches: the layer stays NaN -> 0 */
#define NET_MAX_FEATURES 16

/* OccupancyGridObservation.__init__ arguments (observation.py:286-333) */
typedef struct NetGridCfg {
    int32_t n_features;
    int32_t features[NET_MAX_FEATURES];
    int32_t ranged[NET_MAX_FEATURES]; /* feature has a features_range entry */
    double range_lo[NET_MAX_FEATURES], range_hi[NET_MAX_FEATURES];
    int32_t x_ranged, y_ranged;       /* "x" / "y" in features_range: cell coordinates are un-normalised (:383-400) */
    double x_lo, x_hi, y_lo, y_hi;
    double grid_lo[2], grid_step[2];  /* grid_size[:, 0], grid_step */
    int32_t shape[2];                 /* floor((size[:,1] - size[:,0]) / step) */
    int32_t align_to_vehicle_axes, clip, as_image, observe_intentions;
} NetGridCfg;
/* obs [n_features][shape0][shape1] float32 (as_image: the uint8 values, stored as floats) */
void net_observe_grid(const NetGraph *g, const NetState *s, int V, int ego, const NetGridCfg *gc, float *obs);
/* TimeToCollisionObservation.observe (observation.py:115-152) for observer `ego` with speed index `speed_index`:
 * obs [3][3][horizon * policy_frequency] */
void net_observe_ttc_from(const NetGraph *g, const NetCfg *c, const NetState *s, int V, int ego, int speed_index,
                          float *obs);
/* LidarObservation.observe (observation.py:678-769): obs [cells][2] float32 */
void net_observe_lidar(const NetState *s, int V, int ego, int cells, double maximum_range, int normalize, float *obs);

/* geometry KATs */
void net_lane_local(const NetLane *L, double x, double y, double *s, double *lat);
void net_lane_position(const NetLane *L, double s, double lat, double *x, double *y);
double net_lane_heading_at(const NetLane *L, double s);
int net_closest_lane(const NetGraph *g, double x, double y, double heading);

#ifdef __cplusplus
}
#endif
#endif
