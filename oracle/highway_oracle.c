/*
 * oracle/highway_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See highway_oracle.h for provenance ("PARITY UNPINNED": restatement of the
 * third-party highway-env==1.10.1 pinned by the reference's uv.lock:163-175).
 *
 * Everything here is deliberately naive: fp64, one vehicle object at a time,
 * O(V^2) scans in list order, exactly the sequential semantics SURVEY.md
 * Appendix A describes.  Section markers (A.x) refer to that appendix, the
 * upstream module named beside each function is where the algorithm is
 * published.
 */
#include "highway_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

/* ---- constants (A.3-A.7) ------------------------------------------------- */
static const double LANE_WIDTH = 4.0;
static const double ROAD_LENGTH = 10000.0;
static const double SPEED_LIMIT = 30.0;
static const double VEH_LENGTH = 5.0;
static const double VEH_WIDTH = 2.0;
static const double MAX_SPEED = 40.0;
static const double MIN_SPEED = -40.0;
static const double PERCEPTION_DISTANCE = 200.0; /* 5 * MAX_SPEED */
/* ControlledVehicle */
static const double TAU_HEADING = 0.2, TAU_LATERAL = 0.6, TAU_ACC = 0.6;
static const double MAX_STEERING_ANGLE = PI / 3.0;
/* IDMVehicle */
static const double ACC_MAX = 6.0;
static const double COMFORT_ACC_MAX = 3.0;
static const double COMFORT_ACC_MIN = -5.0;
static const double DISTANCE_WANTED = 10.0; /* 5 + LENGTH */
static const double TIME_WANTED = 1.5;
static const double POLITENESS = 0.0;
static const double LANE_CHANGE_MIN_ACC_GAIN = 0.2;
static const double LANE_CHANGE_MAX_BRAKING_IMPOSED = 2.0;
static const double LANE_CHANGE_DELAY = 1.0;

/* decision kinds (first component of a key) */
enum {
    HW_D_LANE = 1, HW_D_ON_BAND, HW_D_ON_ROAD, HW_D_REACH, HW_D_ORDER, HW_D_NZ_SIGN, HW_D_MOBIL_SAFE, HW_D_MOBIL_GAIN,
    HW_D_ABORT_AHEAD, HW_D_ABORT_GAP, HW_D_SPEED1, HW_D_SPEED_INDEX, HW_D_VMAX, HW_D_PRECHECK, HW_D_SAT_NOW,
    HW_D_SAT_WILL, HW_D_SAT_AXIS, HW_D_OBS_CLOSE, HW_D_OBS_BEHIND, HW_D_OBS_ORDER
};

typedef struct {
    double x, y, heading, speed;
    double target_speed, delta, timer;
    double impact_x, impact_y;
    double act_steer, act_acc; /* the persistent self.action dict */
    int lane, target_lane, crashed, has_impact;
    int is_controlled; /* ControlledVehicle subclass (IDM or MDPVehicle) */
    int is_idm;
} veh_t;

struct hw_env {
    hw_cfg cfg;
    int V;
    veh_t v[HW_MAX_VEHICLES];
    double time;
    int64_t steps;
    uint32_t episode;
    uint32_t obs_draw;
    double min_margin;
    /* test aid: keyed decisions (see dec()) */
    int frame;                 /* simulation frame of the running policy step, HW_FRAME_* outside the frame loop */
    double record_below;       /* decisions decided by less than this are listed in marg[] */
    uint64_t marg[HW_MAX_MARGINAL];
    double marg_val[HW_MAX_MARGINAL];
    int nmarg;
    uint64_t forced[HW_MAX_FORCED];
    int nforced;
    uint8_t slow[HW_MAX_VEHICLES]; /* controlled vehicle acted below 0.5 m/s during the last step */
    double *trace;                 /* test aid: [frames][V][HW_TRACE_FIELDS] per-frame snapshots of hw_step */
};

/* ---- decision bookkeeping (test aid only) --------------------------------
 * Every discrete decision of a step -- a comparison whose outcome selects a branch, an argmin, a
 * sort order -- goes through dec(): `margin` is the signed distance of the decided quantity from
 * its threshold and `res` the outcome the fp64 arithmetic gives.  A decision is named by a key
 * (kind, frame, a, b, c) that does not depend on control flow, so that the SAME logical decision
 * evaluated several times (e.g. "vehicle j is on the band of lane L in frame f", asked once per
 * neighbour query) is one key.  dec() records the smallest |margin| of the step, lists the keys
 * decided by less than record_below, and returns the outcome INVERTED for keys in forced[]: the
 * either-branch parity check of tests/test_env_gpu.py re-runs a step with marginal decisions
 * forced the other way and requires the fp32 kernel to equal one of the outcomes. */
enum { HW_FRAME_PRE = 250, HW_FRAME_END = 251, HW_FRAME_OBS = 252 };
static inline uint64_t dkey(const hw_env *e, int kind, int a, int b, int c)
{
    return ((uint64_t)(kind & 0xff) << 32) | ((uint64_t)(e->frame & 0xff) << 24) | ((uint64_t)(a & 0xff) << 16) |
           ((uint64_t)(b & 0xff) << 8) | (uint64_t)(c & 0xff);
}
#ifdef HW_NO_DECISION_AIDS
/* the CPU-baseline build (bench.py): the plain algorithm, without the test aids' bookkeeping */
#define dec(e, kind, a, b, c, margin, res) (res)
#else
static int dec(hw_env *e, int kind, int a, int b, int c, double margin, int res)
{
    margin = fabs(margin);
    if (margin < e->min_margin) e->min_margin = margin;
    if (margin < e->record_below || e->nforced) {
        uint64_t key = dkey(e, kind, a, b, c);
        if (margin < e->record_below) {
            int q = 0;
            while (q < e->nmarg && e->marg[q] != key) ++q;
            if (q == e->nmarg && e->nmarg < HW_MAX_MARGINAL) { e->marg[q] = key; e->marg_val[q] = margin; e->nmarg++; }
        }
        for (int q = 0; q < e->nforced; ++q)
            if (e->forced[q] == key) return !res;
    }
    return res;
}
#endif
/* strict order of two vehicles along the road, one key per unordered pair */
static int x_before(hw_env *e, int a, int b)
{
#ifdef HW_NO_DECISION_AIDS
    return e->v[a].x < e->v[b].x;
#else
    int lo = a < b ? a : b, hi = a < b ? b : a;
    int lt = e->v[lo].x < e->v[hi].x, gt = e->v[lo].x > e->v[hi].x; /* equal: neither is strictly before */
    int c = dec(e, HW_D_ORDER, lo, hi, 0, e->v[lo].x - e->v[hi].x, lt);
    if (c != lt) gt = !c; /* forced: strictly the other way */
    return a < b ? c : gt;
#endif
}

/* ---- utils.py ------------------------------------------------------------ */
static inline double not_zero(double x)
{
    const double eps = 1e-2;
    if (fabs(x) > eps) return x;
    return x >= 0 ? eps : -eps;
}
static inline double pymod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0 && ((r < 0) != (b < 0))) r += b;
    return r;
}
static inline double wrap_to_pi(double x) { return pymod(x + PI, 2 * PI) - PI; }
static inline double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline double lmap(double v, double x0, double x1, double y0, double y1)
{
    return y0 + (v - x0) * (y1 - y0) / (x1 - x0);
}

/* ---- Philox4x32-10 (Salmon et al., SC'11; public algorithm) -------------- */
void hw_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* 24-bit uniform in [0,1): exactly representable in fp32 and fp64 */
static inline double u01(uint32_t r) { return (double)(r >> 8) * (1.0 / 16777216.0); }

/* ---- road geometry: 4 straight lanes, lane i centred on y = 4 i (A.3) ---- */
static int closest_lane(hw_env *e, int who, double x, double y, double heading)
{
    /* RoadNetwork.get_closest_lane_index with distance_with_heading */
    double best = 0, second = INFINITY;
    int arg = -1, arg2 = -1;
    double ang = fabs(wrap_to_pi(heading - 0.0));
    for (int i = 0; i < e->cfg.lanes_count; ++i) {
        double r = y - LANE_WIDTH * i;
        double d = fabs(r) + fmax(x - ROAD_LENGTH, 0) + fmax(0 - x, 0) + 1.0 * ang;
        if (arg < 0 || d < best) { second = (arg < 0) ? INFINITY : best; arg2 = arg; best = d; arg = i; }
        else if (d < second) { second = d; arg2 = i; }
    }
    /* forced: the runner-up lane */
    if (e->cfg.lanes_count > 1 && dec(e, HW_D_LANE, who, 0, 0, second - best, 0)) return arg2;
    return arg;
}
static int on_lane(hw_env *e, int who, double x, double y, int lane, double margin)
{
    double lat = y - LANE_WIDTH * lane;
    int in = dec(e, margin > 0 ? HW_D_ON_BAND : HW_D_ON_ROAD, who, lane, 0, fabs(lat) - (LANE_WIDTH / 2 + margin),
                 fabs(lat) <= LANE_WIDTH / 2 + margin);
    return in && -VEH_LENGTH <= x && x < ROAD_LENGTH + VEH_LENGTH;
}
static int is_reachable_from(hw_env *e, int who, double x, double y, int lane)
{
    double lat = y - LANE_WIDTH * lane;
    int in = dec(e, HW_D_REACH, who, lane, 0, fabs(lat) - 2 * LANE_WIDTH, fabs(lat) <= 2 * LANE_WIDTH);
    return in && 0 <= x && x < ROAD_LENGTH + VEH_LENGTH;
}

/* ---- Road.neighbour_vehicles (A.6) --------------------------------------- */
static void neighbours(hw_env *e, int self, int lane, int *front, int *rear)
{
    /* upstream: s <= s_v and (s_front is None or s_v <= s_front) -> front; s_v < s and (s_rear is None or
     * s_v > s_rear) -> rear.  Stated through x_before() so that every pairwise order is one keyed decision. */
    int vf = -1, vr = -1;
    for (int j = 0; j < e->V; ++j) {
        if (j == self) continue;
        const veh_t *o = &e->v[j];
        if (!on_lane(e, j, o->x, o->y, lane, 1.0)) continue;
        int j_behind = x_before(e, j, self);                           /* s_v < s */
        if (!j_behind && (vf < 0 || !x_before(e, vf, j))) vf = j;       /* s <= s_v and s_v <= s_front */
        if (j_behind && (vr < 0 || x_before(e, vr, j))) vr = j;         /* s_v < s and s_v > s_rear */
    }
    *front = vf; *rear = vr;
}

/* ---- IDMVehicle.desired_gap / acceleration (A.6) ------------------------- */
static double desired_gap(const veh_t *ego, const veh_t *front)
{
    double ab = -COMFORT_ACC_MAX * COMFORT_ACC_MIN;
    double ce = cos(ego->heading), se = sin(ego->heading);
    double cf = cos(front->heading), sf = sin(front->heading);
    double dv = (ego->speed * ce - front->speed * cf) * ce + (ego->speed * se - front->speed * sf) * se;
    return DISTANCE_WANTED + ego->speed * TIME_WANTED + ego->speed * dv / (2 * sqrt(ab));
}
/* `self` supplies DELTA; `ego` is the vehicle the acceleration is evaluated for */
static double idm_acceleration(const veh_t *self, const veh_t *ego, const veh_t *front)
{
    if (!ego) return 0.0;
    double tgt = ego->is_controlled ? ego->target_speed : 0.0; /* getattr(ego,"target_speed",0) */
    tgt = clipd(tgt, 0, SPEED_LIMIT);
    double a = COMFORT_ACC_MAX * (1 - pow(fmax(ego->speed, 0) / fabs(not_zero(tgt)), self->delta));
    if (front) {
        double d = front->x - ego->x; /* lane_distance_to on a straight lane */
        double g = desired_gap(ego, front) / not_zero(d);
        a -= COMFORT_ACC_MAX * (g * g);
    }
    return a;
}

/* ---- ControlledVehicle.steering_control / speed_control (A.5) ------------ */
static double steering_control(hw_env *e, const veh_t *v, int target_lane)
{
    int who = (int)(v - e->v);
    /* not_zero(speed): the eps switch is continuous, the sign at |speed| < eps is a decision */
    double nzs = not_zero(v->speed);
    if (fabs(v->speed) <= 1e-2 && dec(e, HW_D_NZ_SIGN, who, 0, 0, v->speed, 0)) nzs = -nzs;
    /* test aid: below 0.5 m/s the two divisions by the speed amplify a 1e-6 rounding of y by > 1e3
     * (the controller is ill-conditioned as v -> 0): the vehicle is flagged, the tests widen its tolerances */
    if (fabs(v->speed) < 0.5) e->slow[who] = 1;
    double lat = v->y - LANE_WIDTH * target_lane;
    double lateral_speed_command = -(1.0 / TAU_LATERAL) * lat;
    double heading_command = asin(clipd(lateral_speed_command / nzs, -1, 1));
    double heading_ref = 0.0 + clipd(heading_command, -PI / 4, PI / 4);
    double heading_rate_command = (1.0 / TAU_HEADING) * wrap_to_pi(heading_ref - v->heading);
    double slip = asin(clipd(VEH_LENGTH / 2 / nzs * heading_rate_command, -1, 1));
    double steer = atan(2 * tan(slip));
    return clipd(steer, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
}
static double speed_control(const veh_t *v, double target_speed)
{
    return (1.0 / TAU_ACC) * (target_speed - v->speed);
}

/* ---- IDMVehicle.mobil (A.6) ---------------------------------------------- */
static int mobil(hw_env *e, int self, int cand)
{
    veh_t *me = &e->v[self];
    int np_, nf, op, of;
    neighbours(e, self, cand, &np_, &nf);
    const veh_t *new_prec = np_ >= 0 ? &e->v[np_] : NULL;
    const veh_t *new_foll = nf >= 0 ? &e->v[nf] : NULL;
    double new_following_a = idm_acceleration(me, new_foll, new_prec);
    double new_following_pred_a = idm_acceleration(me, new_foll, me);
    if (dec(e, HW_D_MOBIL_SAFE, self, cand, 0, new_following_pred_a + LANE_CHANGE_MAX_BRAKING_IMPOSED,
            new_following_pred_a < -LANE_CHANGE_MAX_BRAKING_IMPOSED))
        return 0;
    neighbours(e, self, me->lane, &op, &of);
    const veh_t *old_prec = op >= 0 ? &e->v[op] : NULL;
    const veh_t *old_foll = of >= 0 ? &e->v[of] : NULL;
    double self_pred_a = idm_acceleration(me, me, new_prec);
    /* route is None for highway-v0 traffic, so only the acceleration-gain branch runs */
    double self_a = idm_acceleration(me, me, old_prec);
    double old_following_a = idm_acceleration(me, old_foll, me);
    double old_following_pred_a = idm_acceleration(me, old_foll, old_prec);
    double jerk = self_pred_a - self_a +
                  POLITENESS * (new_following_pred_a - new_following_a + old_following_pred_a - old_following_a);
    if (dec(e, HW_D_MOBIL_GAIN, self, cand, 0, jerk - LANE_CHANGE_MIN_ACC_GAIN, jerk < LANE_CHANGE_MIN_ACC_GAIN)) return 0;
    return 1;
}

/* ---- IDMVehicle.change_lane_policy (A.6) --------------------------------- */
static void change_lane_policy(hw_env *e, int self)
{
    veh_t *me = &e->v[self];
    if (me->lane != me->target_lane) {
        for (int j = 0; j < e->V; ++j) {
            const veh_t *o = &e->v[j];
            if (j != self && o->lane != me->target_lane && o->is_controlled &&
                o->target_lane == me->target_lane) {
                double d = o->x - me->x;
                double d_star = desired_gap(me, o);
                if (dec(e, HW_D_ABORT_AHEAD, self, j, 0, d, 0 < d) &&
                    dec(e, HW_D_ABORT_GAP, self, j, 0, d - d_star, d < d_star)) {
                    me->target_lane = me->lane;
                    break;
                }
            }
        }
        return;
    }
    /* utils.do_every; no margin is recorded: the product keeps this timer in fp64 with the
     * same additions, so the comparison is exact (15 x (1/15) == 0.9999999999999999 < 1) */
    if (!(LANE_CHANGE_DELAY < me->timer)) return;
    me->timer = 0;
    int cands[2], nc = 0;
    if (me->lane > 0) cands[nc++] = me->lane - 1;
    if (me->lane < e->cfg.lanes_count - 1) cands[nc++] = me->lane + 1;
    for (int k = 0; k < nc; ++k) {
        if (!is_reachable_from(e, self, me->x, me->y, cands[k])) continue;
        if (dec(e, HW_D_SPEED1, self, 0, 0, fabs(me->speed) - 1, fabs(me->speed) < 1)) continue;
        if (mobil(e, self, cands[k])) me->target_lane = cands[k];
    }
}

/* ---- IDMVehicle.act ------------------------------------------------------ */
static void idm_act(hw_env *e, int self)
{
    veh_t *me = &e->v[self];
    if (me->crashed) return;
    change_lane_policy(e, self);
    double steer = steering_control(e, me, me->target_lane);
    steer = clipd(steer, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
    int f, r;
    neighbours(e, self, me->lane, &f, &r);
    double acc = idm_acceleration(me, me, f >= 0 ? &e->v[f] : NULL);
    if (me->lane != me->target_lane) {
        neighbours(e, self, me->target_lane, &f, &r);
        double acc_t = idm_acceleration(me, me, f >= 0 ? &e->v[f] : NULL);
        acc = fmin(acc, acc_t);
    }
    acc = clipd(acc, -ACC_MAX, ACC_MAX);
    me->act_steer = steer;
    me->act_acc = acc;
}

/* ---- MDPVehicle (DiscreteMetaAction ego; secondary mode) ------ */
static int speed_to_index(hw_env *e, int who, double speed)
{
    double x = (speed - 20.0) / (30.0 - 20.0);
    double r = nearbyint(x * 2.0); /* np.round: half to even (default FE_TONEAREST) */
    if (e) { /* forced: the other neighbouring integer */
        double f = x * 2.0 - floor(x * 2.0);
        if (dec(e, HW_D_SPEED_INDEX, who, 0, 0, f - 0.5, 0)) r = (r > x * 2.0) ? r - 1 : r + 1;
    }
    return (int)clipd(r, 0, 2);
}
static void controlled_act(hw_env *e, int self, int action)
{
    /* actions {0:LANE_LEFT,1:IDLE,2:LANE_RIGHT,3:FASTER,4:SLOWER}; -1 == None */
    veh_t *me = &e->v[self];
    if (action == 3 || action == 4) {
        int idx = speed_to_index(e, self, me->speed) + (action == 3 ? 1 : -1);
        idx = (int)clipd(idx, 0, 2);
        me->target_speed = 20.0 + 5.0 * idx; /* np.linspace(20, 30, 3) */
    } else if (action == 2 || action == 0) {
        int t = (int)clipd(me->target_lane + (action == 2 ? 1 : -1), 0, e->cfg.lanes_count - 1);
        if (is_reachable_from(e, self, me->x, me->y, t)) me->target_lane = t;
    }
    double steer = steering_control(e, me, me->target_lane);
    me->act_steer = clipd(steer, -MAX_STEERING_ANGLE, MAX_STEERING_ANGLE);
    me->act_acc = speed_control(me, me->target_speed);
}

/* ---- Vehicle.step / clip_actions / on_state_update (A.4) ----------------- */
static void vehicle_step(hw_env *e, int idx, double dt)
{
    veh_t *v = &e->v[idx];
    if (v->is_idm) v->timer += dt;
    if (v->crashed) { v->act_steer = 0; v->act_acc = -1.0 * v->speed; }
    if (dec(e, HW_D_VMAX, idx, 0, 0, v->speed - MAX_SPEED, v->speed > MAX_SPEED))
        v->act_acc = fmin(v->act_acc, 1.0 * (MAX_SPEED - v->speed));
    else if (dec(e, HW_D_VMAX, idx, 1, 0, v->speed - MIN_SPEED, v->speed < MIN_SPEED))
        v->act_acc = fmax(v->act_acc, 1.0 * (MIN_SPEED - v->speed));
    double beta = atan(1.0 / 2 * tan(v->act_steer));
    double vx = v->speed * cos(v->heading + beta);
    double vy = v->speed * sin(v->heading + beta);
    v->x += vx * dt;
    v->y += vy * dt;
    if (v->has_impact) {
        v->x += v->impact_x; v->y += v->impact_y;
        v->crashed = 1; v->has_impact = 0; v->impact_x = v->impact_y = 0;
    }
    v->heading += v->speed * sin(beta) / (VEH_LENGTH / 2) * dt;
    v->speed += v->act_acc * dt;
    v->lane = closest_lane(e, idx, v->x, v->y, v->heading);
}

/* ---- utils.are_polygons_intersecting + RoadObject.handle_collisions (A.7) - */
static void polygon(const veh_t *v, double p[5][2])
{
    const double lx[4] = {-VEH_LENGTH / 2, -VEH_LENGTH / 2, +VEH_LENGTH / 2, +VEH_LENGTH / 2};
    const double ly[4] = {-VEH_WIDTH / 2, +VEH_WIDTH / 2, +VEH_WIDTH / 2, -VEH_WIDTH / 2};
    double c = cos(v->heading), s = sin(v->heading);
    for (int k = 0; k < 4; ++k) {
        p[k][0] = c * lx[k] - s * ly[k] + v->x;
        p[k][1] = s * lx[k] + c * ly[k] + v->y;
    }
    p[4][0] = p[0][0]; p[4][1] = p[0][1];
}
static void project(const double p[5][2], const double n[2], double *mn, double *mx)
{
    *mn = *mx = 0;
    for (int k = 0; k < 5; ++k) {
        double d = p[k][0] * n[0] + p[k][1] * n[1];
        if (k == 0 || d < *mn) *mn = d;
        if (k == 0 || d > *mx) *mx = d;
    }
}
static inline double interval_distance(double min_a, double max_a, double min_b, double max_b)
{
    return min_a < min_b ? min_b - max_a : min_a - max_b;
}
static void handle_collisions(hw_env *e, int ia, int ib, double dt)
{
    veh_t *A = &e->v[ia], *B = &e->v[ib];
    double dx = B->x - A->x, dy = B->y - A->y;
    double diag = sqrt(VEH_LENGTH * VEH_LENGTH + VEH_WIDTH * VEH_WIDTH);
    double lim = (diag + diag) / 2 + A->speed * dt;
    double dist = sqrt(dx * dx + dy * dy);
    if (dec(e, HW_D_PRECHECK, ia, ib, 0, dist - lim, dist > lim)) return;
    double a[5][2], b[5][2];
    polygon(A, a); polygon(B, b);
    double da[2] = {A->speed * cos(A->heading) * dt, A->speed * sin(A->heading) * dt};
    double db[2] = {B->speed * cos(B->heading) * dt, B->speed * sin(B->heading) * dt};
    int intersecting = 1, will_intersect = 1;
    double min_distance = INFINITY, axis[2] = {0, 0};
    double seen_d[8], seen_ax[8][2];
    int nseen = 0;
    for (int poly = 0; poly < 2; ++poly) {
        const double(*P)[2] = poly == 0 ? a : b;
        for (int k = 0; k < 4; ++k) {
            double n[2] = {-P[k + 1][1] + P[k][1], P[k + 1][0] - P[k][0]};
            double nn = sqrt(n[0] * n[0] + n[1] * n[1]);
            n[0] /= nn; n[1] /= nn;
            double min_a, max_a, min_b, max_b;
            project(a, n, &min_a, &max_a);
            project(b, n, &min_b, &max_b);
            double sd = interval_distance(min_a, max_a, min_b, max_b);
            int edge = 2 * poly + (k & 1); /* edges k and k + 2 of a rectangle are opposite normals of ONE axis: one key */
            if (dec(e, HW_D_SAT_NOW, ia, ib, edge, sd, sd > 0)) intersecting = 0;
            double vp = n[0] * (da[0] - db[0]) + n[1] * (da[1] - db[1]);
            if (vp < 0) min_a += vp; else max_a += vp;
            double distance = interval_distance(min_a, max_a, min_b, max_b);
            if (dec(e, HW_D_SAT_WILL, ia, ib, edge, distance, distance > 0)) will_intersect = 0;
            if (!intersecting && !will_intersect) break;
            {
                double cx = 0, cy = 0;
                for (int q = 0; q < 4; ++q) { cx += a[q][0] - b[q][0]; cy += a[q][1] - b[q][1]; }
                double dd = (cx / 4) * n[0] + (cy / 4) * n[1];
                double sx = dd > 0 ? n[0] : -n[0], sy = dd > 0 ? n[1] : -n[1];
                seen_d[nseen] = fabs(distance); seen_ax[nseen][0] = sx; seen_ax[nseen][1] = sy; ++nseen;
                /* the axis choice matters only through the impact vector: the margin is recorded when
                 * the two axes differ (n and -n of one rectangle give the same vector) */
                int differs = fabs(sx - axis[0]) + fabs(sy - axis[1]) > 1e-9;
                int wins = fabs(distance) < min_distance;
                if (differs && isfinite(min_distance))
                    wins = dec(e, HW_D_SAT_AXIS, ia, ib, edge, fabs(distance) - min_distance, wins);
                if (wins) {
                    min_distance = fabs(distance);
                    axis[0] = sx; axis[1] = sy;
                }
            }
        }
    }
    (void)seen_d; (void)seen_ax;
    if (will_intersect) {
        double tx = min_distance * axis[0], ty = min_distance * axis[1];
        A->impact_x = tx / 2; A->impact_y = ty / 2; A->has_impact = 1;
        B->impact_x = -tx / 2; B->impact_y = -ty / 2; B->has_impact = 1;
    }
    if (intersecting) { A->crashed = 1; B->crashed = 1; }
}

/* ---- AbstractEnv.step / _simulate (A.2, A.11) ---------------------------- */
static int ego_on_road(hw_env *e)
{
    const veh_t *ego = &e->v[0];
    return on_lane(e, 0, ego->x, ego->y, ego->lane, 0.0);
}

void hw_step(hw_env *e, const float *action, double *reward, int32_t *terminated, int32_t *truncated)
{
    const hw_cfg *c = &e->cfg;
    e->min_margin = INFINITY;
    e->nmarg = 0;
    memset(e->slow, 0, sizeof(e->slow));
    e->frame = HW_FRAME_PRE;
    e->time += 1.0 / c->policy_frequency;
    int frames = c->simulation_frequency / c->policy_frequency;
    double dt = 1.0 / c->simulation_frequency;
    veh_t *ego = &e->v[0];
    for (int frame = 0; frame < frames; ++frame) {
        e->frame = frame;
        if (e->steps % frames == 0) {
            if (c->ego_mode == 0) {
                /* ContinuousAction.get_action: float32 arithmetic on the np.float32 action */
                float a0 = action[0], a1 = action[1];
                a0 = a0 < -1.f ? -1.f : (a0 > 1.f ? 1.f : a0);
                a1 = a1 < -1.f ? -1.f : (a1 > 1.f ? 1.f : a1);
                float acc = -5.0f + (a0 - (-1.0f)) * 10.0f / 2.0f;
                float qpi = (float)(PI / 4);
                float st = (-qpi) + (a1 - (-1.0f)) * (float)(PI / 4 - (-PI / 4)) / 2.0f;
                ego->act_acc = (double)acc;
                ego->act_steer = (double)st;
            } else {
                controlled_act(e, 0, (int)action[0]);
            }
        }
        /* Road.act(): list order */
        for (int i = 0; i < e->V; ++i) {
            if (e->v[i].is_idm) idm_act(e, i);
            else if (c->ego_mode == 1) controlled_act(e, i, -1);
            /* plain Vehicle.act(None): keeps its action */
        }
        /* Road.step(dt) */
        for (int i = 0; i < e->V; ++i) vehicle_step(e, i, dt);
        for (int i = 0; i < e->V; ++i)
            for (int j = i + 1; j < e->V; ++j) handle_collisions(e, i, j, dt);
        e->steps += 1;
        if (e->trace) {
            double *t = e->trace + (size_t)frame * e->V * HW_TRACE_FIELDS;
            for (int i = 0; i < e->V; ++i, t += HW_TRACE_FIELDS) {
                const veh_t *v = &e->v[i];
                t[0] = v->x; t[1] = v->y; t[2] = v->speed; t[3] = v->heading; t[4] = v->impact_x; t[5] = v->impact_y;
                t[6] = (double)((uint32_t)v->lane | ((uint32_t)v->target_lane << 8) | ((uint32_t)(v->crashed != 0) << 16) |
                                ((uint32_t)(v->has_impact != 0) << 17));
            }
        }
    }
    e->frame = HW_FRAME_END;
    /* HighwayEnv._rewards / _reward (A.9) */
    int lane = ego->is_controlled ? ego->target_lane : ego->lane;
    double forward_speed = ego->speed * cos(ego->heading);
    double scaled = lmap(forward_speed, c->reward_speed_lo, c->reward_speed_hi, 0, 1);
    int nlanes = c->lanes_count;
    double r_coll = ego->crashed ? 1.0 : 0.0;
    double r_lane = (double)lane / (double)((nlanes - 1) > 1 ? (nlanes - 1) : 1);
    double r_speed = clipd(scaled, 0, 1);
    int onr = ego_on_road(e);
    double r = c->collision_reward * r_coll + c->right_lane_reward * r_lane + c->high_speed_reward * r_speed +
               0.0 * (double)onr;
    if (c->normalize_reward)
        r = lmap(r, c->collision_reward, c->high_speed_reward + c->right_lane_reward, 0, 1);
    r *= (double)onr;
    *reward = r;
    *terminated = (ego->crashed || (c->offroad_terminal && !onr)) ? 1 : 0;
    *truncated = e->time >= c->duration ? 1 : 0;
}

/* ---- KinematicObservation.observe + Road.close_objects_to (A.8) ---------- */
static double feature_value(const veh_t *v, int code)
{
    switch (code) {
    case HW_F_PRESENCE: return 1.0;
    case HW_F_X: return v->x;
    case HW_F_Y: return v->y;
    case HW_F_VX: return v->speed * cos(v->heading);
    case HW_F_VY: return v->speed * sin(v->heading);
    case HW_F_HEADING: return v->heading;
    case HW_F_COS_H: return cos(v->heading);
    case HW_F_SIN_H: return sin(v->heading);
    }
    return 0.0;
}
/* sort key |x_p - x_ego| of p strictly greater than that of q (p listed before q: p < q) */
static int obs_after(hw_env *e, int p, int q)
{
    double kp = fabs(e->v[p].x - e->v[0].x), kq = fabs(e->v[q].x - e->v[0].x);
    return dec(e, HW_D_OBS_ORDER, p, q, 0, kp - kq, kp > kq);
}
void hw_observe(const hw_env *ce, float *obs, const int32_t *perm, int32_t *row_vehicle)
{
    hw_env *e = (hw_env *)ce; /* decision bookkeeping only */
    const hw_cfg *c = &e->cfg;
    e->frame = HW_FRAME_OBS;
    int N = c->obs_vehicles, F = c->obs_nfeat;
    const veh_t *ego = &e->v[0];
    int cand[HW_MAX_VEHICLES], nc = 0;
    for (int j = 1; j < e->V; ++j) {
        const veh_t *o = &e->v[j];
        double dx = o->x - ego->x, dy = o->y - ego->y;
        double dist = sqrt(dx * dx + dy * dy);
        if (!dec(e, HW_D_OBS_CLOSE, j, 0, 0, dist - PERCEPTION_DISTANCE, dist < PERCEPTION_DISTANCE)) continue;
        if (!c->obs_see_behind) {
            if (!dec(e, HW_D_OBS_BEHIND, j, 0, 0, dx + 2 * VEH_LENGTH, -2 * VEH_LENGTH < dx)) continue;
        }
        cand[nc++] = j;
    }
    if (c->obs_sorted) { /* Python sorted(): stable, key |lane_distance_to| */
        for (int a = 1; a < nc; ++a) { /* insertion sort: stable; one keyed decision per compared pair */
            int cur = cand[a];
            int b = a - 1;
            while (b >= 0 && obs_after(e, cand[b], cur)) { cand[b + 1] = cand[b]; --b; }
            cand[b + 1] = cur;
        }
    }
    if (nc > N - 1) nc = N - 1;
    double table[HW_MAX_VEHICLES][HW_MAX_FEATURES];
    int rowveh[HW_MAX_VEHICLES];
    for (int r = 0; r < N; ++r) {
        rowveh[r] = -1;
        for (int f = 0; f < F; ++f) table[r][f] = 0.0;
    }
    for (int r = 0; r <= nc && r < N; ++r) {
        const veh_t *v = r == 0 ? ego : &e->v[cand[r - 1]];
        rowveh[r] = r == 0 ? 0 : cand[r - 1];
        for (int f = 0; f < F; ++f) {
            int code = c->obs_feat[f];
            double val = feature_value(v, code);
            if (r > 0 && !c->obs_absolute &&
                (code == HW_F_X || code == HW_F_Y || code == HW_F_VX || code == HW_F_VY))
                val -= feature_value(ego, code);
            if (c->obs_normalize && c->obs_has_range[f]) {
                val = lmap(val, c->obs_lo[f], c->obs_hi[f], -1, 1);
                if (c->obs_clip) val = clipd(val, -1, 1);
            }
            table[r][f] = val;
        }
    }
    for (int r = 0; r < N; ++r) {
        int dst = r;
        if (r > 0 && !c->obs_sorted && perm) dst = 1 + perm[r - 1];
        for (int f = 0; f < F; ++f) obs[dst * F + f] = (float)table[r][f];
        if (row_vehicle) row_vehicle[dst] = rowveh[r];
    }
}

void hw_shuffle_perm(uint64_t seed, uint64_t env_id, uint32_t draw, int32_t n, int32_t *perm)
{
    /* rank of a random 32-bit key per row (ties by index): a uniform permutation */
    uint32_t keys[HW_MAX_VEHICLES];
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ 0xA5A5A5A5u};
    for (int blk = 0; blk * 4 < n; ++blk) {
        uint32_t ctr[4] = {(uint32_t)blk, draw, (uint32_t)env_id, (uint32_t)(env_id >> 32)}, out[4];
        hw_philox4x32_10(ctr, key, out);
        for (int q = 0; q < 4 && blk * 4 + q < n; ++q) keys[blk * 4 + q] = out[q];
    }
    for (int i = 0; i < n; ++i) {
        int rank = 0;
        for (int j = 0; j < n; ++j)
            if (keys[j] < keys[i] || (keys[j] == keys[i] && j < i)) ++rank;
        perm[i] = rank;
    }
}

/* ---- HighwayEnv._create_vehicles / Vehicle.create_random (A.3) ------------ */
void hw_reset(hw_env *e, uint64_t seed, uint64_t env_id, uint32_t episode)
{
    const hw_cfg *c = &e->cfg;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    double gap = exp(-5.0 / 40.0 * c->lanes_count);
    double xprev = 0;
    e->time = 0; e->steps = 0; e->episode = episode;
    e->min_margin = INFINITY;
    for (int k = 0; k < e->V; ++k) {
        uint32_t ctr[4] = {(uint32_t)k, episode, (uint32_t)env_id, (uint32_t)(env_id >> 32)}, r[4];
        hw_philox4x32_10(ctr, key, r);
        veh_t *v = &e->v[k];
        memset(v, 0, sizeof(*v));
        int lane = (int)(u01(r[0]) * c->lanes_count);
        double speed, spacing;
        if (k == 0) {
            if (c->initial_lane_id >= 0) lane = c->initial_lane_id;
            speed = 25.0; spacing = c->ego_spacing;
        } else {
            speed = 0.7 * SPEED_LIMIT + (0.8 * SPEED_LIMIT - 0.7 * SPEED_LIMIT) * u01(r[1]);
            spacing = 1.0 / c->vehicles_density;
        }
        double offset = spacing * (12.0 + 1.0 * speed) * gap;
        double x0 = k == 0 ? 3 * offset : xprev; /* max x so far == the previous spawn */
        x0 += offset * (0.9 + (1.1 - 0.9) * u01(r[2]));
        xprev = x0;
        v->x = x0; v->y = LANE_WIDTH * lane; v->heading = 0; v->speed = speed;
        v->lane = lane; v->target_lane = lane; v->target_speed = speed;
        v->is_idm = k > 0;
        v->is_controlled = k > 0 || c->ego_mode == 1;
        v->delta = k > 0 ? 3.5 + (4.5 - 3.5) * u01(r[3]) : 4.0;
        v->timer = k > 0 ? pymod((v->x + v->y) * PI, LANE_CHANGE_DELAY) : 0.0;
        if (k == 0 && c->ego_mode == 1) v->target_speed = 20.0 + 5.0 * speed_to_index(NULL, 0, speed);
    }
}

hw_env *hw_create(const hw_cfg *cfg)
{
    if (cfg->vehicles_count + 1 > HW_MAX_VEHICLES || cfg->obs_vehicles > HW_MAX_VEHICLES ||
        cfg->obs_nfeat > HW_MAX_FEATURES)
        return NULL;
    hw_env *e = (hw_env *)calloc(1, sizeof(hw_env));
    e->cfg = *cfg;
    e->V = cfg->vehicles_count + 1;
    hw_reset(e, 0, 0, 0);
    return e;
}
void hw_destroy(hw_env *e) { free(e); }
int hw_num_vehicles(const hw_env *e) { return e->V; }
double hw_last_min_margin(const hw_env *e) { return e->min_margin; }
void hw_record_margin(hw_env *e, double below) { e->record_below = below; }
int32_t hw_marginal_keys(const hw_env *e, uint64_t *keys, double *margins, int32_t max)
{
    int n = e->nmarg < max ? e->nmarg : max;
    for (int q = 0; q < n; ++q) { keys[q] = e->marg[q]; if (margins) margins[q] = e->marg_val[q]; }
    return e->nmarg;
}
int32_t hw_force_decisions(hw_env *e, const uint64_t *keys, int32_t n)
{
    if (n < 0 || n > HW_MAX_FORCED) return -1;
    for (int q = 0; q < n; ++q) e->forced[q] = keys[q];
    e->nforced = n;
    return 0;
}
void hw_slow_vehicles(const hw_env *e, uint8_t *out) { memcpy(out, e->slow, (size_t)e->V); }
void hw_set_trace(hw_env *e, double *buf) { e->trace = buf; }

void hw_get_state(const hw_env *e, hw_state *s)
{
    memset(s, 0, sizeof(*s));
    for (int k = 0; k < e->V; ++k) {
        const veh_t *v = &e->v[k];
        s->x[k] = v->x; s->y[k] = v->y; s->heading[k] = v->heading; s->speed[k] = v->speed;
        s->target_speed[k] = v->target_speed; s->delta[k] = v->delta; s->timer[k] = v->timer;
        s->impact_x[k] = v->impact_x; s->impact_y[k] = v->impact_y;
        s->lane[k] = v->lane; s->target_lane[k] = v->target_lane;
        s->crashed[k] = v->crashed; s->has_impact[k] = v->has_impact;
    }
    s->time = e->time; s->steps = e->steps;
}
void hw_set_state(hw_env *e, const hw_state *s)
{
    for (int k = 0; k < e->V; ++k) {
        veh_t *v = &e->v[k];
        v->x = s->x[k]; v->y = s->y[k]; v->heading = s->heading[k]; v->speed = s->speed[k];
        v->target_speed = s->target_speed[k]; v->delta = s->delta[k]; v->timer = s->timer[k];
        v->impact_x = s->impact_x[k]; v->impact_y = s->impact_y[k];
        v->lane = s->lane[k]; v->target_lane = s->target_lane[k];
        v->crashed = s->crashed[k]; v->has_impact = s->has_impact[k];
        v->is_idm = k > 0;
        v->is_controlled = k > 0 || e->cfg.ego_mode == 1;
        v->act_steer = 0; v->act_acc = 0;
    }
    e->time = s->time; e->steps = s->steps;
}

/* ---- CPU-baseline batch driver: independent envs, OpenMP over envs -------- */
void hw_batch_step(hw_env **envs, int32_t n, const float *actions, float *obs, float *reward,
                   uint8_t *terminated, uint8_t *truncated, uint64_t seed, int32_t autoreset,
                   int32_t nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int i = 0; i < n; ++i) {
        hw_env *e = envs[i];
        double r; int32_t te, tr;
        hw_step(e, actions + 2 * i, &r, &te, &tr);
        reward[i] = (float)r; terminated[i] = (uint8_t)te; truncated[i] = (uint8_t)tr;
        if (autoreset && (te || tr)) hw_reset(e, seed, (uint64_t)i, e->episode + 1);
        int N = e->cfg.obs_vehicles, F = e->cfg.obs_nfeat;
        int32_t perm[HW_MAX_VEHICLES];
        const int32_t *pp = NULL;
        if (!e->cfg.obs_sorted) {
            hw_shuffle_perm(seed, (uint64_t)i, e->obs_draw, N - 1, perm);
            pp = perm;
        }
        e->obs_draw += 1;
        hw_observe(e, obs + (size_t)i * N * F, pp, NULL);
    }
}
