/*
 * oracle/highway_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (fp64, sequential) restatement of the highway-v0 simulator that the
 * reference drives through gymnasium (reference call sites:
 * experiments/wrappers.py:76-80, training/routine.py:18,24,127,134).
 *
 * The simulator itself is the third-party package highway-env==1.10.1
 * (reference uv.lock:163-175) on gymnasium==1.1.1 (uv.lock:140-149); neither
 * is vendored under /root/reference nor installable here.  This file restates
 * the published algorithm of that package as exercised by
 * config/base_config.py:5-39 (SURVEY.md Appendix A is the spec).
 *
 * PARITY UNPINNED: the reference's own tests never touch the simulator, and
 * upstream highway-env could not be run in this build, so the simulator half
 * of the oracle is checked only for self-consistency (known-answer vectors
 * derived by hand from the published formulas, see tests/test_oracle_cpu.py).
 * profiles/r02_highway_env_probe.txt records that the GPU box has no
 * highway_env / gymnasium either.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef HIGHWAY_ORACLE_H
#define HIGHWAY_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HW_MAX_VEHICLES 128
#define HW_MAX_FEATURES 8
#define HW_MAX_MARGINAL 1024 /* marginal decisions listed per step */
#define HW_MAX_FORCED 64   /* decisions that can be forced the other way at once */

/* feature codes for the Kinematics observation (highway-env to_dict keys) */
enum {
    HW_F_PRESENCE = 0, HW_F_X = 1, HW_F_Y = 2, HW_F_VX = 3, HW_F_VY = 4,
    HW_F_HEADING = 5, HW_F_COS_H = 6, HW_F_SIN_H = 7
};

typedef struct hw_cfg {
    int32_t lanes_count;          /* 4   */
    int32_t vehicles_count;       /* 50 other vehicles (V = vehicles_count + 1) */
    int32_t simulation_frequency; /* 15  */
    int32_t policy_frequency;     /* 1   */
    int32_t initial_lane_id;      /* -1 == None */
    int32_t ego_mode;             /* 0 ContinuousAction, 1 DiscreteMetaAction (MDPVehicle) */
    int32_t normalize_reward;     /* 1   */
    int32_t offroad_terminal;     /* 0   */
    double  duration;             /* 40  */
    double  ego_spacing;          /* 2   */
    double  vehicles_density;     /* 2 in the reference config */
    double  collision_reward;     /* -1  */
    double  right_lane_reward;    /* 0.1 */
    double  high_speed_reward;    /* 0.4 */
    double  reward_speed_lo;      /* 20  */
    double  reward_speed_hi;      /* 30  */
    /* observation (KinematicObservation kwargs) */
    int32_t obs_vehicles;         /* N rows, 15 */
    int32_t obs_nfeat;            /* F */
    int32_t obs_feat[HW_MAX_FEATURES];
    int32_t obs_has_range[HW_MAX_FEATURES];
    double  obs_lo[HW_MAX_FEATURES];
    double  obs_hi[HW_MAX_FEATURES];
    int32_t obs_normalize;        /* 1 */
    int32_t obs_clip;             /* 1 */
    int32_t obs_absolute;         /* 0 */
    int32_t obs_sorted;           /* 1 == order "sorted", 0 == "shuffled" */
    int32_t obs_see_behind;       /* 0 */
    int32_t _pad;
} hw_cfg;

/* One env's full simulator state, SoA over the V vehicles (ego is index 0). */
typedef struct hw_state {
    double  x[HW_MAX_VEHICLES];
    double  y[HW_MAX_VEHICLES];
    double  heading[HW_MAX_VEHICLES];
    double  speed[HW_MAX_VEHICLES];
    double  target_speed[HW_MAX_VEHICLES];
    double  delta[HW_MAX_VEHICLES];      /* IDM exponent, randomize_behavior() */
    double  timer[HW_MAX_VEHICLES];      /* IDM lane-change timer */
    double  impact_x[HW_MAX_VEHICLES];
    double  impact_y[HW_MAX_VEHICLES];
    int32_t lane[HW_MAX_VEHICLES];
    int32_t target_lane[HW_MAX_VEHICLES];
    int32_t crashed[HW_MAX_VEHICLES];
    int32_t has_impact[HW_MAX_VEHICLES];
    double  time;                        /* env.time   */
    int64_t steps;                       /* env.steps  */
} hw_state;

typedef struct hw_env hw_env;

hw_env *hw_create(const hw_cfg *cfg);
void    hw_destroy(hw_env *env);
int     hw_num_vehicles(const hw_env *env);

/* counter-based spawn: Philox4x32-10, key = seed, counter = (vehicle, episode, env_id) */
void hw_reset(hw_env *env, uint64_t seed, uint64_t env_id, uint32_t episode);
void hw_get_state(const hw_env *env, hw_state *out);
void hw_set_state(hw_env *env, const hw_state *in);

/* One policy step.  action: 2 floats (continuous) or action[0] = meta-action id. */
void hw_step(hw_env *env, const float *action, double *reward, int32_t *terminated,
             int32_t *truncated);

/* Kinematics observation. obs: N*F float32.  perm (nullable): N-1 destinations for
 * the shuffled order (row 1+k of the unshuffled table goes to row 1+perm[k]).
 * row_vehicle (nullable): vehicle index shown in every output row, -1 for padding. */
void hw_observe(const hw_env *env, float *obs, const int32_t *perm, int32_t *row_vehicle);

/* the permutation the product path draws for (seed, env_id, draw) */
void hw_shuffle_perm(uint64_t seed, uint64_t env_id, uint32_t draw, int32_t n, int32_t *perm);

/* raw generator, for the known-answer test */
void hw_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* batch helpers for the CPU baseline (OpenMP over envs) */
void hw_batch_step(hw_env **envs, int32_t n, const float *actions, float *obs,
                   float *reward, uint8_t *terminated, uint8_t *truncated,
                   uint64_t seed, int32_t autoreset, int32_t nthreads);

/* smallest margin by which any discrete decision taken during the last hw_step()
 * was decided (lane argmin, timer, MOBIL thresholds, SAT separations ...) */
double hw_last_min_margin(const hw_env *env);

/* ---- either-branch parity aid.  Every discrete decision has a control-flow independent key
 * (kind << 32 | frame << 24 | a << 16 | b << 8 | c).  hw_record_margin: decisions of the following
 * steps / observations that are decided by less than `below` are listed (hw_marginal_keys returns
 * how many there were; the list restarts with every hw_step).  hw_force_decisions: the listed keys
 * are decided the OTHER way from now on (n = 0 clears).  hw_slow_vehicles: controlled vehicles that
 * acted below 0.5 m/s in the last step (ill-conditioned steering: tolerances are widened). */
void    hw_record_margin(hw_env *env, double below);
int32_t hw_marginal_keys(const hw_env *env, uint64_t *keys, double *margins, int32_t max);
int32_t hw_force_decisions(hw_env *env, const uint64_t *keys, int32_t n);
void    hw_slow_vehicles(const hw_env *env, uint8_t *out);
/* while buf is non-NULL, hw_step writes every vehicle's state at the end of every simulation frame:
 * buf[frames][V][HW_TRACE_FIELDS] = (x, y, speed, heading, impact_x, impact_y, lane | target_lane<<8 | crashed<<16 |
 * has_impact<<17) */
#define HW_TRACE_FIELDS 7
void    hw_set_trace(hw_env *env, double *buf);

#ifdef __cplusplus
}
#endif
#endif
