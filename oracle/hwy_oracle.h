/*
 * hwy_oracle.h — CPU restatement of the HighwayEnv hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This header and hwy_oracle.c are the parity ORACLE for the CUDA path in
 * highwayenv_b200/csrc.  Nothing in the product (highwayenv_b200/, include/) may
 * include, link or call it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * It is a scalar, sequential, per-vehicle restatement that follows the reference's own
 * loop structure (list order, Gauss-Seidel target-lane updates, last-writer-wins
 * impacts).  Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  Parity is PINNED: tests/test_oracle_golden.py checks it against
 * golden trajectories produced by the unmodified Python reference (oracle/gen_golden.py,
 * fixtures in tests/golden/), and tests/test_oracle_live.py against the live reference
 * when /root/reference is present.
 */
#ifndef HWY_ORACLE_H
#define HWY_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LANES 8
#define ORC_MAX_TARGET_SPEEDS 8

/* vehicle kinds */
#define ORC_MAX_OBS_FEATURES 16
#define ORC_KIND_IDM 0      /* highway_env/vehicle/behavior.py:12  IDMVehicle */
#define ORC_KIND_MDP 1      /* highway_env/vehicle/controller.py:256 MDPVehicle (DiscreteMetaAction ego) */
#define ORC_KIND_VEHICLE 2  /* highway_env/vehicle/kinematics.py:13 Vehicle (ContinuousAction ego) */

/* Scenario configuration of the straight multi-lane highway family
 * (highway_env/envs/highway_env.py:25-53,162-175; abstract.py:102-125). */
typedef struct OrcHighwayCfg {
    int32_t lanes_count;
    int32_t n_vehicles; /* controlled (1) + vehicles_count */
    int32_t simulation_frequency;
    int32_t policy_frequency;
    int32_t action_type; /* 0 DiscreteMetaAction, 1 ContinuousAction */
    int32_t others_check_collisions; /* 0: highway-fast (highway_env.py:177-182), 1: highway */
    int32_t normalize_reward;
    int32_t offroad_terminal;
    int32_t obs_vehicles_count; /* KinematicObservation.vehicles_count */
    int32_t obs_see_behind;
    int32_t obs_absolute;
    int32_t obs_normalize;
    int32_t obs_clip;
    int32_t n_target_speeds;
    int32_t initial_lane_id; /* -1 = None */
    int32_t act_clip;        /* ContinuousAction.clip */
    double duration;
    double lane_length;  /* road.py:295 (10000) */
    double lane_width;   /* lane.py:16 (4) */
    double speed_limit;  /* highway_env.py:63 (30) */
    double target_speeds[ORC_MAX_TARGET_SPEEDS];
    double collision_reward, right_lane_reward, high_speed_reward;
    double reward_speed_lo, reward_speed_hi;
    double acc_lo, acc_hi, steer_lo, steer_hi; /* action.py:82-86 */
    double ego_spacing, vehicles_density, ego_speed;
    double spawn_exp; /* exp(-5/40*lanes) evaluated by numpy on the host (kinematics.py:95) */
    /* IDM / MOBIL parameters (behavior.py:21-46) */
    double acc_max, comfort_acc_max, comfort_acc_min, distance_wanted, time_wanted;
    double politeness, lane_change_min_acc_gain, lane_change_max_braking_imposed, lane_change_delay;
    double delta_lo, delta_hi; /* DELTA_RANGE */
    double perception_distance; /* abstract.py:56 */
    /* KinematicObservation.features / features_range (observation.py:160-232): 0 = the default
     * (presence, x, y, vx, vy) with default ranges; else the Vehicle.to_dict keys (kinematics.py:240-254)
     * by code: 0 presence 1 x 2 y 3 vx 4 vy 5 heading 6 cos_h 7 sin_h 8 cos_d 9 sin_d 10 long_off
     * 11 lat_off 12 ang_off; a column is normalised iff obs_feature_ranged */
    int32_t obs_n_features, _pad_obs;
    int32_t obs_feature[ORC_MAX_OBS_FEATURES], obs_feature_ranged[ORC_MAX_OBS_FEATURES];
    double obs_feature_lo[ORC_MAX_OBS_FEATURES], obs_feature_hi[ORC_MAX_OBS_FEATURES];
} OrcHighwayCfg;

/* One env's state, structure of arrays over vehicles (list order = road.vehicles). */
typedef struct OrcHighwayState {
    double *x, *y, *heading, *speed;
    double *target_speed, *timer, *delta;
    double *impact_x, *impact_y;
    int32_t *lane, *target_lane;
    int32_t *kind;             /* ORC_KIND_* */
    int32_t *crashed;
    int32_t *has_impact;
    int32_t *check_collisions;
    int32_t *speed_index;      /* [1] ego MDPVehicle.speed_index */
    double *time;              /* [1] env.time */
} OrcHighwayState;

/* numpy Generator(PCG64) bit-exact restatement (numpy/random/src/pcg64, distributions.c). */
typedef struct OrcPcg64 {
    uint64_t state_hi, state_lo, inc_hi, inc_lo;
    uint32_t has_uint32, uinteger;
} OrcPcg64;

uint64_t orc_pcg64_next64(OrcPcg64 *g);
uint32_t orc_pcg64_next32(OrcPcg64 *g);
double orc_pcg64_double(OrcPcg64 *g);
double orc_rng_uniform(OrcPcg64 *g, double lo, double hi);
int64_t orc_rng_choice(OrcPcg64 *g, int64_t n); /* Generator.choice(n) / integers(n) */

/* geometry known-answer helpers (utils.py) */
double orc_wrap_to_pi(double x);
double orc_not_zero(double x);
int orc_rotated_rectangles_intersect(double c1x, double c1y, double l1, double w1, double a1,
                                     double c2x, double c2y, double l2, double w2, double a2);
void orc_polygons_intersecting(const double a[5][2], const double b[5][2], double dax, double day,
                               double dbx, double dby, int *intersecting, int *will_intersect,
                               double trans[2]);

/* highway family */
void orc_highway_reset(const OrcHighwayCfg *cfg, OrcPcg64 *rng, OrcHighwayState *st);
void orc_highway_observe(const OrcHighwayCfg *cfg, const OrcHighwayState *st, float *obs);
/* action: int (DiscreteMetaAction) in action_i, or float32 action_f[2] (ContinuousAction;
 * the Box dtype — the reference's lmap then runs in float32, see continuous_act). */
void orc_highway_step(const OrcHighwayCfg *cfg, OrcHighwayState *st, int action_i,
                      const float *action_f, float *obs, double *reward, int32_t *terminated,
                      int32_t *truncated);
/* Road.act() + Road.step(dt) only, `substeps` times with no ego action. */
void orc_highway_substeps(const OrcHighwayCfg *cfg, OrcHighwayState *st, int substeps);

/* Batched driver over SoA [n_envs][V] buffers with SameStep auto-reset; `threads` host
 * threads (contiguous env ranges).  Used for large-scale parity and the CPU baseline. */
typedef struct OrcBatch {
    int32_t n_envs;
    double *x, *y, *heading, *speed, *target_speed, *timer, *delta, *impact_x, *impact_y;
    int32_t *lane, *target_lane, *kind, *crashed, *has_impact, *check_collisions;
    int32_t *speed_index; /* [n_envs] */
    double *time;         /* [n_envs] */
    OrcPcg64 *rng;        /* [n_envs] */
} OrcBatch;

void orc_highway_reset_batch(const OrcHighwayCfg *cfg, OrcBatch *b, const uint8_t *mask,
                             float *obs, int threads);
void orc_highway_step_batch(const OrcHighwayCfg *cfg, OrcBatch *b, const int32_t *action_i,
                            const float *action_f, float *obs, double *reward,
                            uint8_t *terminated, uint8_t *truncated, int autoreset, int threads);

#ifdef __cplusplus
}
#endif
#endif
