// hrp_env.cu -- the fused highway-v0 step / observe / reset kernels (sm_100a).
//
// One warp owns one env for the whole policy step: the 15 simulation frames of
// HighwayEnv._simulate (reference call site training/routine.py:134; algorithm: highway-env
// 1.10.1) run out of shared memory, then reward/termination, the
// optional in-place respawn, the Kinematics observation and the RoPE/Dist/Rank embedding
// (experiments/rope_embed.py:64-74, dist_embed.py:76-96, rank_embed.py:45-51) are produced by
// the same warp before the state goes back to HBM once.
//
// Lane l of the warp owns vehicles l and l+32 (V <= 64) in registers.  What other vehicles
// need to see of a vehicle lives in a per-warp shared-memory block.  Instead of the
// reference's O(V) scan per neighbour query, vehicles are kept ranked by x (fp64 keys,
// repaired by odd-even transposition each frame) and every lane band has a 64-bit occupancy
// mask in rank order, so "vehicle in front / behind on lane L" is two bit operations.
//
// Numerics: x and the IDM lane-change timer are fp64 (ordering and the 1.0 < timer test are
// decided exactly as the fp64 reference decides them); everything else is fp32.
#include <math.h>

#include "hrp_internal.cuh"

namespace {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kLaneW = 4.0f;
constexpr float kInvTwoSqrtAB = 0.12909944487358055f;  // 1 / (2 sqrt(3*5))
constexpr float kTanBetaMax = 0.86602540378443865f;    // tan(pi/3) / 2
constexpr float kDiag = 5.385164807134504f;            // sqrt(5^2 + 2^2)

struct __align__(16) WarpS {
    double x[HRP_VS];    // absolute longitudinal position
    double key[HRP_VS];  // scratch: observation sort keys
    float xr[HRP_VS];    // x - xref, fp32 working copy
    float y[HRP_VS], v[HRP_VS], ch[HRP_VS], sh[HRP_VS], h[HRP_VS], ts[HRP_VS], delta[HRP_VS];
    float impx[HRP_VS], impy[HRP_VS];
    int impkey[HRP_VS];
    float rowd[HRP_VS];  // scratch: per-row normalised distance of the embedding
    ull band[HRP_MAX_LANES];
    unsigned char lane[HRP_VS], tl_old[HRP_VS], tl_new[HRP_VS], order[HRP_VS], rank[HRP_VS];
    unsigned char crash[HRP_VS], queue[HRP_VS], res[2 * HRP_VS], perm[HRP_VS];
    uint32_t skey[HRP_VS];
    int rowveh[HRP_MAX_OBS_ROWS];
    float obs[HRP_MAX_OBS_ROWS * HRP_MAX_FEATURES];
};

struct Veh {  // what the owning lane keeps in registers
    double x, timer;
    float y, h, v, ts, delta, impx, impy, ch, sh;
    float acc, tb;  // persistent action: acceleration and tan(beta)
    int lane, tlane;
    bool crashed, has_impact;
};

__device__ __forceinline__ float nzf(float x) { return fabsf(x) > 1e-2f ? x : (x >= 0.f ? 1e-2f : -1e-2f); }
__device__ __forceinline__ float clipf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float wrap_to_pi(float x)
{
    if (x >= -kPi && x < kPi) return x;
    return x - 2.f * kPi * floorf((x + kPi) / (2.f * kPi));
}
// sin and cos of a heading: traffic headings are a few tenths of a radian, where the Taylor polynomials
// (|error| < 3e-10 for |x| <= 0.5) are exact to fp32 rounding; anything larger takes the library path
__device__ __forceinline__ void sincos_heading(float x, float *s, float *c)
{
    if (fabsf(x) <= 0.5f) {
        float x2 = x * x;
        float ps = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.7557319e-6f, -1.9841270e-4f), 8.3333333e-3f), -1.6666667e-1f), 1.f);
        float pc = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.f);
        *s = x * ps;
        *c = pc;
    } else {
        sincosf(x, s, c);
    }
}
// first set bit above / last set bit below rank p of a 64-slot occupancy mask, on the two 32-bit halves (the
// 64-bit shift / ffs / clz sequences the compiler emits for the one-line versions cost twice as many instructions)
__device__ __forceinline__ int slot_front(ull mask, int p)
{
    const uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
    const uint32_t mlo = p < 31 ? lo & (0xFFFFFFFEu << p) : 0u;                                   // bits p+1 .. 31
    const uint32_t mhi = p < 32 ? hi : (p < 63 ? hi & (0xFFFFFFFEu << (p - 32)) : 0u);             // bits max(p+1, 32) .. 63
    return mlo ? __ffs((int)mlo) - 1 : (mhi ? 31 + __ffs((int)mhi) : -1);
}
__device__ __forceinline__ int slot_rear(ull mask, int p)
{
    const uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
    const uint32_t mhi = p > 32 ? hi & ((1u << (p - 32)) - 1u) : 0u;                               // bits 32 .. p-1
    const uint32_t mlo = p >= 32 ? lo : (p > 0 ? lo & ((1u << p) - 1u) : 0u);                      // bits 0 .. min(p, 32)-1
    return mhi ? 63 - __clz((int)mhi) : (mlo ? 31 - __clz((int)mlo) : -1);
}
__device__ __forceinline__ int closest_lane(float y, int lanes)
{
    // argmin_i |y - 4 i| with the first minimum winning a tie (RoadNetwork.get_closest_lane_index):
    // ceil(y / 4 - 1/2), clamped.  y / 4 and the subtraction of 0.5 are exact in binary floating point
    // wherever the result can change the answer.
    int i = (int)ceilf(y * 0.25f - 0.5f);
    return max(0, min(lanes - 1, i));
}

// IDMVehicle.desired_gap: e follows f
__device__ __forceinline__ float desired_gap(const WarpS &S, int e, int f)
{
    float ve = S.v[e], che = S.ch[e], she = S.sh[e];
    float dvx = ve * che - S.v[f] * S.ch[f];
    float dvy = ve * she - S.v[f] * S.sh[f];
    float dv = dvx * che + dvy * she;
    return 10.f + ve * 1.5f + ve * dv * kInvTwoSqrtAB;
}
// COMFORT_ACC_MAX * (1 - (max(v,0)/|not_zero(v0)|)^delta); v0 already clipped to [0, 30]
__device__ __forceinline__ float idm_free(float v, float v0, float delta)
{
    float r = __fdividef(fmaxf(v, 0.f), fabsf(nzf(v0)));  // <= 2 ulp; the result goes through exp2/log2 anyway
    float p = r > 0.f ? exp2f(delta * __log2f(r)) : 0.f;
    return 3.f * (1.f - p);
}
// COMFORT_ACC_MAX * (desired_gap / not_zero(d))^2
__device__ __forceinline__ float idm_interaction(const WarpS &S, int e, int f)
{
    float d = S.xr[f] - S.xr[e];
    float g = __fdividef(desired_gap(S, e, f), nzf(d));
    return 3.f * g * g;
}

// the same for the calling lane's own vehicle, whose state is in registers
__device__ __forceinline__ float idm_interaction_own(const WarpS &S, float xr, float ve, float che, float she, int f)
{
    float dvx = ve * che - S.v[f] * S.ch[f];
    float dvy = ve * she - S.v[f] * S.sh[f];
    float gap = 10.f + ve * 1.5f + ve * (dvx * che + dvy * she) * kInvTwoSqrtAB;
    float g = __fdividef(gap, nzf(S.xr[f] - xr));
    return 3.f * g * g;
}

// IDMVehicle.mobil for vehicle i and the candidate on `side` (0: lane-1, 1: lane+1).
// Returns candidate lane + 1, or 0.  The free-road term of self_pred_a - self_a cancels.
__device__ int mobil_item(const WarpS &S, const EnvDev &P, int i, int side)
{
    int li = S.lane[i];
    int c = side == 0 ? li - 1 : li + 1;
    if (c < 0 || c >= P.lanes) return 0;
    double xi = S.x[i];
    if (!(fabsf(S.y[i] - kLaneW * c) <= 2.f * kLaneW && xi >= 0.0 && xi < 10005.0)) return 0;
    if (fabsf(S.v[i]) < 1.f) return 0;
    int p = S.rank[i];
    ull mc = S.band[c];
    int sf = slot_front(mc, p), sr = slot_rear(mc, p);
    if (sr >= 0) {
        int nf = S.order[sr];
        bool ctrl = nf > 0 || P.ego_mode == 1;
        float v0 = ctrl ? clipf(S.ts[nf], 0.f, 30.f) : 0.f;
        float a = idm_free(S.v[nf], v0, S.delta[i]) - idm_interaction(S, nf, i);
        if (a < -2.f) return 0;
    }
    float pred = sf >= 0 ? idm_interaction(S, i, S.order[sf]) : 0.f;
    int so = slot_front(S.band[li], p);
    float cur = so >= 0 ? idm_interaction(S, i, S.order[so]) : 0.f;
    float jerk = cur - pred;
    return jerk >= 0.2f ? c + 1 : 0;
}

// RoadObject.handle_collisions for list-ordered pair (i < j): spherical pre-check, then the
// separating-axis test of utils.are_polygons_intersecting.  The reference walks the 8 edge
// normals (n0, n1, -n0, -n1 of a, then of b); n and -n give the same separation flags, and the
// same |distance| unless one projected interval contains the other, so each of the 4 distinct
// normals is evaluated once and both distance variants feed the minimum in the reference order.
// Returns bit0 = intersecting, bit1 = will_intersect (+ translation).
__device__ int collide_pair(const WarpS &S, int i, int j, float dt, float &tx, float &ty)
{
    float dx = S.xr[j] - S.xr[i], dy = S.y[j] - S.y[i];
    float lim = kDiag + S.v[i] * dt;
    if (lim < 0.f || dx * dx + dy * dy > lim * lim) return 0;
    float uax = S.ch[i], uay = S.sh[i], ubx = S.ch[j], uby = S.sh[j];
    float cax = S.xr[i], cay = S.y[i], cbx = S.xr[j], cby = S.y[j];
    float rdx = (S.v[i] * uax - S.v[j] * ubx) * dt, rdy = (S.v[i] * uay - S.v[j] * uby) * dt;
    {
        // cheap exit for the common near miss (vehicles side by side on adjacent lanes): if a's lateral
        // axis separates the rectangles now AND after the displacement, both flags end up false whatever
        // the other axes say, which is the reference's "no contact" result
        float nx = -uay, ny = uax;
        float pa = cax * nx + cay * ny, pb = cbx * nx + cby * ny;
        float rb = 2.5f * fabsf(ubx * nx + uby * ny) + fabsf(-uby * nx + ubx * ny);
        float gap = fabsf(pa - pb) - (1.f + rb);
        if (gap > 0.f && gap - fabsf(nx * rdx + ny * rdy) > 0.f) return 0;
    }
    bool inter = true, will = true;
    float mind = INFINITY, ax = 0.f, ay = 0.f;
    float dneg[2], nnx[2], nny[2], ddn[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // reference edge order: -u_a, +w_a, (+u_a, -w_a), -u_b, +w_b, (+u_b, -w_b)
        float nx, ny;
        if (k == 0) { nx = -uax; ny = -uay; }
        else if (k == 1) { nx = -uay; ny = uax; }
        else if (k == 2) { nx = -ubx; ny = -uby; }
        else { nx = -uby; ny = ubx; }
        float pa = cax * nx + cay * ny, pb = cbx * nx + cby * ny;
        float ra = 2.5f * fabsf(uax * nx + uay * ny) + fabsf(-uay * nx + uax * ny);
        float rb = 2.5f * fabsf(ubx * nx + uby * ny) + fabsf(-uby * nx + ubx * ny);
        float min_a = pa - ra, max_a = pa + ra, min_b = pb - rb, max_b = pb + rb;
        float sd = min_a < min_b ? min_b - max_a : min_a - max_b;
        if (sd > 0.f) inter = false;
        float vp = nx * rdx + ny * rdy;
        if (vp < 0.f) min_a += vp; else max_a += vp;
        float d1 = min_b - max_a, d2 = min_a - max_b;
        float dist = min_a < min_b ? d1 : d2;
        if (dist > 0.f) will = false;
        if (!inter && !will) return 0;
        float dd = (cax - cbx) * nx + (cay - cby) * ny;
        if (fabsf(dist) < mind) {
            mind = fabsf(dist);
            if (dd > 0.f) { ax = nx; ay = ny; } else { ax = -nx; ay = -ny; }
        }
        dneg[k & 1] = max_a > max_b ? d2 : d1;  // the same edge seen through the opposite normal
        nnx[k & 1] = nx; nny[k & 1] = ny; ddn[k & 1] = dd;
        if (k & 1) {
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (fabsf(dneg[t]) < mind) {
                    mind = fabsf(dneg[t]);
                    if (ddn[t] < 0.f) { ax = -nnx[t]; ay = -nny[t]; } else { ax = nnx[t]; ay = nny[t]; }
                }
        }
    }
    tx = mind * ax; ty = mind * ay;
    return (inter ? 1 : 0) | (will ? 2 : 0);
}

// ---------------------------------------------------------------------------------------
// rank bookkeeping
__device__ void rank_full(WarpS &S, int V, int lane)
{
    // rank = number of vehicles ordered before (x, list index); O(V) per vehicle, once per launch
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        if (k < V) {
            double xk = S.x[k];
            int r = 0;
            for (int j = 0; j < V; ++j) {
                double xj = S.x[j];
                r += (xj < xk || (xj == xk && j < k)) ? 1 : 0;
            }
            S.rank[k] = (unsigned char)r;
            S.order[r] = (unsigned char)k;
        }
    }
    __syncwarp();
}
__device__ void rank_repair(WarpS &S, int V, int lane)
{
    // fp32 pre-check on the ego-relative copies: an inversion (true gap <= 0) shows up as an fp32 gap below the
    // margin (|xr| < 4096 m: two roundings <= 5e-4 m), so "every gap >= 4e-3" proves the order without fp64 loads
    {
        bool sus = false;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            int s = lane + 32 * q;
            if (s + 1 < V) sus |= S.xr[S.order[s + 1]] - S.xr[S.order[s]] < 4e-3f;
        }
        if (!__any_sync(HRP_FULL, sus)) return;   // ranks are unchanged, nothing to rewrite
    }
    for (;;) {
        bool inv = false;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            int s = lane + 32 * q;
            if (s + 1 < V) inv |= S.x[S.order[s]] > S.x[S.order[s + 1]];
        }
        if (!__any_sync(HRP_FULL, inv)) break;
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            int s = 2 * lane + par;
            if (s + 1 < V) {
                int a = S.order[s], b = S.order[s + 1];
                if (S.x[a] > S.x[b]) { S.order[s] = (unsigned char)b; S.order[s + 1] = (unsigned char)a; }
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int s = lane + 32 * q;
        if (s < V) S.rank[S.order[s]] = (unsigned char)s;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// state movement HBM <-> registers/shared
__device__ void publish(WarpS &S, const Veh &u, int k, double xref)
{
    S.x[k] = u.x; S.xr[k] = (float)(u.x - xref);
    S.y[k] = u.y; S.v[k] = u.v; S.ch[k] = u.ch; S.sh[k] = u.sh; S.h[k] = u.h;
    S.ts[k] = u.ts; S.delta[k] = u.delta;
    S.lane[k] = (unsigned char)u.lane; S.tl_old[k] = S.tl_new[k] = (unsigned char)u.tlane;
    S.impkey[k] = -1; S.crash[k] = 0;
}
__device__ void load_env(const EnvDev &P, WarpS &S, Veh (&u)[2], int e, int lane, double &xref)
{
    size_t base = (size_t)e * HRP_VS;
    xref = P.x[base];  // ego x at the start of the step
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        if (k < P.V) {
            Veh &w = u[q];
            w.x = P.x[base + k]; w.timer = P.timer[base + k];
            w.y = P.y[base + k]; w.h = P.heading[base + k]; w.v = P.speed[base + k];
            w.ts = P.tspeed[base + k]; w.delta = P.delta[base + k];
            w.impx = P.impx[base + k]; w.impy = P.impy[base + k];
            uint32_t f = P.flags[base + k];
            w.lane = f & 0xff; w.tlane = (f >> 8) & 0xff;
            w.crashed = (f >> 16) & 1; w.has_impact = (f >> 17) & 1;
            sincos_heading(w.h, &w.sh, &w.ch);
            w.acc = 0.f; w.tb = 0.f;
            publish(S, w, k, xref);
        } else {
            u[q] = Veh{};  // empty slot: the frame code runs on it with selects, its results are never stored
            u[q].ch = 1.f;
        }
    }
    __syncwarp();
}
__device__ void store_env(const EnvDev &P, const Veh (&u)[2], int e, int lane)
{
    size_t base = (size_t)e * HRP_VS;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        if (k < P.V) {
            const Veh &w = u[q];
            P.x[base + k] = w.x; P.timer[base + k] = w.timer;
            P.y[base + k] = w.y; P.heading[base + k] = w.h; P.speed[base + k] = w.v;
            P.tspeed[base + k] = w.ts; P.delta[base + k] = w.delta;
            P.impx[base + k] = w.impx; P.impy[base + k] = w.impy;
            P.flags[base + k] = (uint32_t)w.lane | ((uint32_t)w.tlane << 8) |
                                ((uint32_t)w.crashed << 16) | ((uint32_t)w.has_impact << 17);
        }
    }
}

// ---------------------------------------------------------------------------------------
// HighwayEnv._create_vehicles / Vehicle.create_random with Philox draws.
__device__ __forceinline__ float u01f(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
__device__ int speed_to_index(float speed)
{
    float x = (speed - 20.f) / 10.f;
    return (int)clipf(rintf(x * 2.f), 0.f, 2.f);
}
__device__ void spawn_env(const EnvDev &P, WarpS &S, Veh (&u)[2], int e, int lane, uint32_t episode,
                          double &xref)
{
    ull gid = P.env_id_base + (ull)e;
    double val[2] = {0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        Veh &w = u[q];
        if (k < P.V) {
            uint32_t r[4];
            hrp_philox((uint32_t)k, episode, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)P.seed,
                       (uint32_t)(P.seed >> 32), r);
            int ln = (int)((double)u01f(r[0]) * (double)P.lanes);
            double speed, spacing;
            if (k == 0) {
                if (P.initial_lane >= 0) ln = P.initial_lane;
                speed = 25.0; spacing = P.ego_spacing;
            } else {
                speed = 21.0 + 3.0 * (double)u01f(r[1]);
                spacing = P.inv_density;
            }
            double offset = spacing * (12.0 + speed) * P.gap_factor;
            double inc = offset * (0.9 + (1.1 - 0.9) * (double)u01f(r[2]));
            val[q] = k == 0 ? 3.0 * offset + inc : inc;
            w.y = kLaneW * ln; w.h = 0.f; w.v = (float)speed; w.ts = (float)speed;
            w.lane = ln; w.tlane = ln;
            w.delta = k > 0 ? 3.5f + u01f(r[3]) : 4.0f;
            w.impx = w.impy = 0.f; w.crashed = false; w.has_impact = false;
            w.ch = 1.f; w.sh = 0.f; w.acc = 0.f; w.tb = 0.f;
            if (k == 0 && P.ego_mode == 1) w.ts = 20.f + 5.f * speed_to_index(w.v);
        }
    }
    // x_k = x_{k-1} + inc_k: inclusive scan over list order (fp64)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double t = __shfl_up_sync(HRP_FULL, val[q], d);
            if (lane >= d) val[q] += t;
        }
    }
    double tot0 = __shfl_sync(HRP_FULL, val[0], 31);
    val[1] += tot0;
    xref = __shfl_sync(HRP_FULL, val[0], 0);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        if (k < P.V) {
            Veh &w = u[q];
            w.x = val[q];
            double t = (w.x + (double)w.y) * 3.14159265358979323846;
            w.timer = k > 0 ? t - floor(t) : 0.0;  // python: (x + y) * pi % 1.0, operands >= 0
            publish(S, w, k, xref);
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// KinematicObservation.observe + embedding wrapper, warp-cooperative.
__device__ double feature_value(const WarpS &S, int code, int k)
{
    switch (code) {
    case HRP_F_PRESENCE: return 1.0;
    case HRP_F_X: return S.x[k];
    case HRP_F_Y: return (double)S.y[k];
    case HRP_F_VX: return (double)S.v[k] * (double)S.ch[k];
    case HRP_F_VY: return (double)S.v[k] * (double)S.sh[k];
    case HRP_F_HEADING: return (double)S.h[k];
    case HRP_F_COS_H: return (double)S.ch[k];
    case HRP_F_SIN_H: return (double)S.sh[k];
    }
    return 0.0;
}
__device__ void write_row(const EnvDev &P, WarpS &S, int row, int k)
{
    S.rowveh[row] = k;
    for (int f = 0; f < P.F; ++f) {
        int code = P.feat[f];
        double val = feature_value(S, code, k);
        if (k != 0 && !P.absolute && code >= HRP_F_X && code <= HRP_F_VY) val -= feature_value(S, code, 0);
        if (P.normalize && P.has_range[f]) {
            val = -1.0 + (val - P.lo[f]) * 2.0 / (P.hi[f] - P.lo[f]);
            if (P.clip) val = fmin(fmax(val, -1.0), 1.0);
        }
        S.obs[row * P.F + f] = (float)val;
    }
}

// the embedding epilogue on an [N, F] table in shared memory -> out[N, Fout] in global memory
__device__ void embed_store(int kind, int edim, int use_euclid, int ego_idx, float max_dist,
                            const float *__restrict__ table, const float *tab, float *rowd, int N,
                            int F, int Fout, int lane, float *__restrict__ out,
                            const float *__restrict__ dist_override = nullptr)
{
    if (dist_override) {  // RotaryEmbedWrapper._apply_rope(obs, dist_norm): caller-supplied distances
        for (int r = lane; r < N; r += 32) rowd[r] = dist_override[r];
        __syncwarp();
    } else if (kind == HRP_EMBED_ROPE || kind == HRP_EMBED_DIST) {
        // rel = obs[:, :2] - obs[ego, :2]; dist = clip(norm(rel) / max_dist, 0, 1), all float32
        float ex = tab[ego_idx * F + 0], ey = F > 1 ? tab[ego_idx * F + 1] : 0.f;
        for (int r = lane; r < N; r += 32) {
            float rx = __fsub_rn(tab[r * F + 0], ex);
            float d;
            if (use_euclid) {
                float ry = __fsub_rn(tab[r * F + 1], ey);
                d = __fsqrt_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)));
            } else {
                d = fabsf(rx);
            }
            d = __fdiv_rn(d, max_dist);
            rowd[r] = fminf(fmaxf(d, 0.f), 1.f);
        }
        __syncwarp();
    }
    const float two_pi = 6.2831853071795862f;  // float32(2*pi), as numpy's weak python scalar
    int total = N * Fout;
    for (int idx = lane; idx < total; idx += 32) {
        int r = idx / Fout, c = idx - r * Fout;
        float o;
        if (kind == HRP_EMBED_ROPE && c < edim) {
            int p = c >> 1;
            float theta = __fmul_rn(__fmul_rn(two_pi, rowd[r]), table[p]);
            float sn, cs;
            sincosf(theta, &sn, &cs);
            float a = tab[r * F + 2 * p], b = tab[r * F + 2 * p + 1];
            o = (c & 1) ? __fadd_rn(__fmul_rn(a, sn), __fmul_rn(b, cs))
                        : __fsub_rn(__fmul_rn(a, cs), __fmul_rn(b, sn));
        } else if (c < F) {
            o = tab[r * F + c];
        } else if (kind == HRP_EMBED_DIST) {
            int j = c - F, half = edim >> 1;
            float fr = table[j < half ? j : j - half];
            float ang = __fmul_rn(__fmul_rn(two_pi, rowd[r]), fr);
            o = j < half ? sinf(ang) : cosf(ang);
        } else {  // HRP_EMBED_RANK: tanh(table) precomputed by the host
            o = table[r * edim + (c - F)];
        }
        out[idx] = o;
    }
}

__device__ void observe_env(const EnvDev &P, WarpS &S, int e, int lane, float *__restrict__ obs,
                            const int32_t *__restrict__ perm_in, int32_t *__restrict__ row_vehicle,
                            uint32_t draw)
{
    const int V = P.V, N = P.N, F = P.F;
    for (int i = lane; i < N * F; i += 32) S.obs[i] = 0.f;
    for (int i = lane; i < N; i += 32) S.rowveh[i] = -1;
    // Road.close_objects_to: list order, ||p - p_ego|| < 200 and -10 < dx unless see_behind
    double ex = S.x[0];
    float ey = S.y[0];
    bool close[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        close[q] = false;
        if (k >= 1 && k < V) {
            double dx = S.x[k] - ex, dy = (double)S.y[k] - (double)ey;
            close[q] = (dx * dx + dy * dy < 200.0 * 200.0) && (P.see_behind || dx > -10.0);
            S.key[k] = close[q] ? fabs(dx) : INFINITY;
        }
    }
    ull cm = (ull)__ballot_sync(HRP_FULL, close[0]) | ((ull)__ballot_sync(HRP_FULL, close[1]) << 32);
    // row permutation for order == "shuffled": injected, or rank of Philox keys
    if (!P.sorted && N > 1) {
        if (perm_in) {
            for (int i = lane; i < N - 1; i += 32) S.perm[i] = (unsigned char)perm_in[(size_t)e * (N - 1) + i];
        } else {
            ull gid = P.env_id_base + (ull)e;
            for (int blk = lane; blk * 4 < N - 1; blk += 32) {
                uint32_t r[4];
                hrp_philox((uint32_t)blk, draw, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)P.seed,
                           (uint32_t)(P.seed >> 32) ^ 0xA5A5A5A5u, r);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (blk * 4 + t < N - 1) S.skey[blk * 4 + t] = r[t];
            }
            __syncwarp();
            for (int i = lane; i < N - 1; i += 32) {
                uint32_t ki = S.skey[i];
                int r = 0;
                for (int j = 0; j < N - 1; ++j) {
                    uint32_t kj = S.skey[j];
                    r += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
                }
                S.perm[i] = (unsigned char)r;
            }
        }
    }
    __syncwarp();
    if (lane == 0) write_row(P, S, 0, 0);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        if (close[q]) {
            int r;
            if (P.sorted) {  // stable sort by |dx|: count candidates ordered before k
                double kk = S.key[k];
                r = 0;
                ull m = cm;
                while (m) {
                    int j = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    double kj = S.key[j];
                    r += (kj < kk || (kj == kk && j < k)) ? 1 : 0;
                }
            } else {
                r = __popcll(cm & ((1ull << k) - 1ull));
            }
            if (r < N - 1) {
                int row = 1 + (P.sorted ? r : (int)S.perm[r]);
                write_row(P, S, row, k);
            }
        }
    }
    __syncwarp();
    float *out = obs + (size_t)e * N * P.Fout;
    embed_store(P.embed_kind, P.embed_dim, P.use_euclid, P.ego_idx, P.max_dist, P.table, S.obs, S.rowd, N,
                F, P.Fout, lane, out);
    if (row_vehicle)
        for (int i = lane; i < N; i += 32) row_vehicle[(size_t)e * N + i] = S.rowveh[i];
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// One simulation frame: Road.act() then Road.step(dt)
__device__ void simulate_frame(const EnvDev &P, WarpS &S, Veh (&u)[2], int lane, double xref)
{
    const int V = P.V;
    const float dt = P.dt;
    const bool ego_ctrl = P.ego_mode == 1;

    // ---- band occupancy masks in rank order: |y - 4L| <= 3 (on_lane with margin 1)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int s = lane + 32 * q;
        float yy = 0.f;
        bool inx = false;
        if (s < V) {
            int veh = S.order[s];
            yy = S.y[veh];
            double xx = S.x[veh];
            inx = xx >= -5.0 && xx < 10005.0;
        }
        for (int L = 0; L < P.lanes; ++L) {
            unsigned b = __ballot_sync(HRP_FULL, inx && fabsf(yy - kLaneW * L) <= 3.f);
            if (lane == 0) reinterpret_cast<unsigned *>(&S.band[L])[q] = b;
        }
    }
    __syncwarp();

    // ---- IDMVehicle.change_lane_policy, part 1: MOBIL decisions of vehicles whose timer fired
    bool fire[2], mid[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        Veh &w = u[q];
        bool idm = k > 0 && k < V && !w.crashed;
        mid[q] = idm && w.lane != w.tlane;
        fire[q] = idm && w.lane == w.tlane && 1.0 < w.timer;
        if (fire[q]) w.timer = 0.0;
    }
    unsigned fm0 = __ballot_sync(HRP_FULL, fire[0]), fm1 = __ballot_sync(HRP_FULL, fire[1]);
    if (fm0 | fm1) {
        unsigned lt = (1u << lane) - 1u;
        int pos[2] = {__popc(fm0 & lt), __popc(fm0) + __popc(fm1 & lt)};
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (fire[q]) S.queue[pos[q]] = (unsigned char)(lane + 32 * q);
        __syncwarp();
        int nitems = 2 * (__popc(fm0) + __popc(fm1));
        for (int t = lane; t < nitems; t += 32)
            S.res[t] = (unsigned char)mobil_item(S, P, S.queue[t >> 1], t & 1);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (fire[q]) {
                int r0 = S.res[2 * pos[q]], r1 = S.res[2 * pos[q] + 1];
                if (r0) u[q].tlane = r0 - 1;
                if (r1) u[q].tlane = r1 - 1;  // a later candidate overwrites an earlier one
                S.tl_new[lane + 32 * q] = (unsigned char)u[q].tlane;
            }
        __syncwarp();
    }

    // ---- part 2: vehicles already changing lane abort if someone ahead targets the same lane.
    // List order matters: vehicle k sees this frame's decisions of j < k, last frame's of j > k.
    ull mm = (ull)__ballot_sync(HRP_FULL, mid[0]) | ((ull)__ballot_sync(HRP_FULL, mid[1]) << 32);
    while (mm) {
        int k = __ffsll((long long)mm) - 1;
        mm &= mm - 1;
        int T = S.tl_new[k];
        float xk = S.xr[k];
        bool hit = false;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            int j = lane + 32 * q;
            if (j < V && j != k && (j > 0 || ego_ctrl) && S.lane[j] != T) {
                int tj = j < k ? S.tl_new[j] : S.tl_old[j];
                if (tj == T) {
                    float d = S.xr[j] - xk;
                    if (d > 0.f && d < desired_gap(S, k, j)) hit = true;
                }
            }
        }
        if (__any_sync(HRP_FULL, hit)) {
            if (lane == (k & 31)) {
                if (k >> 5) u[1].tlane = u[1].lane; else u[0].tlane = u[0].lane;
                S.tl_new[k] = S.lane[k];
            }
        }
        __syncwarp();
    }

    // ---- controls: steering_control + IDM acceleration (or the ego's own action).  Straight-line code with
    // selects: every lane runs the same instructions whether its slot holds a vehicle, a crashed one or none.
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int k = lane + 32 * q;
        Veh &w = u[q];
        const bool valid = k < V;
        const bool act = valid && !w.crashed && (k > 0 || ego_ctrl);  // ContinuousAction ego keeps its action dict
        // ControlledVehicle.steering_control(target_lane) -> tan(beta) without leaving the tangent
        const float lat = w.y - kLaneW * w.tlane;
        const float lsc = -(1.f / 0.6f) * lat;
        const float rv = __fdividef(1.f, nzf(w.v));
        // clip(asin(clip(c, -1, 1)), -pi/4, pi/4) == asin(clip(c, -sin(pi/4), sin(pi/4))): asin is monotone
        const float hc = asinf(clipf(lsc * rv, -0.70710678118654752f, 0.70710678118654752f));
        const float href = clipf(hc, -kPi / 4.f, kPi / 4.f);
        const float hrc = 5.f * wrap_to_pi(href - w.h);
        const float ss = clipf(2.5f * rv * hrc, -1.f, 1.f);  // sin(slip)
        const float tslip = ss * rsqrtf(fmaxf(1.f - ss * ss, 0.f));  // tan(slip); +-inf at |ss| == 1
        const float tb_new = clipf(tslip, -kTanBetaMax, kTanBetaMax);  // tan(beta) = clip(2 tan slip, +-tan(pi/3)) / 2
        // IDMVehicle.acceleration against the front vehicle of the own lane and of the target lane
        const int p = valid ? (int)S.rank[k] : 0;
        const float a0 = idm_free(w.v, clipf(w.ts, 0.f, 30.f), w.delta);
        const float xr_own = (float)(w.x - xref);
        const int sf = slot_front(S.band[w.lane], p), st = slot_front(S.band[w.tlane], p);
        const float i1 = idm_interaction_own(S, xr_own, w.v, w.ch, w.sh, S.order[max(sf, 0)]);
        const float i2 = idm_interaction_own(S, xr_own, w.v, w.ch, w.sh, S.order[max(st, 0)]);
        float acc = a0 - (sf >= 0 ? i1 : 0.f);
        const float acc_t = a0 - (st >= 0 ? i2 : 0.f);
        acc = w.lane != w.tlane ? fminf(acc, acc_t) : acc;
        const float acc_idm = clipf(acc, -6.f, 6.f);
        const float acc_mdp = (1.f / 0.6f) * (w.ts - w.v);  // MDPVehicle: speed_control
        w.tb = act ? tb_new : w.tb;
        w.acc = act ? (k > 0 ? acc_idm : acc_mdp) : w.acc;
    }
    __syncwarp();  // every read of the frame-start state is done

    // ---- Vehicle.step(dt): clip_actions, kinematic bicycle, pending impact, lane re-assignment
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int k = lane + 32 * q;
        Veh &w = u[q];
        w.timer += k > 0 ? P.dt64 : 0.0;
        w.tb = w.crashed ? 0.f : w.tb;
        w.acc = w.crashed ? -1.0f * w.v : w.acc;
        w.acc = w.v > 40.f ? fminf(w.acc, 40.f - w.v) : (w.v < -40.f ? fmaxf(w.acc, -40.f - w.v) : w.acc);
        const float cb = rsqrtf(1.f + w.tb * w.tb), sb = w.tb * cb;
        const float c = w.ch * cb - w.sh * sb, s = w.sh * cb + w.ch * sb;
        w.x += (double)(w.v * c * dt);
        w.y += w.v * s * dt;
        w.x += w.has_impact ? (double)w.impx : 0.0;   // pending impact of the previous frame's collision
        w.y += w.has_impact ? w.impy : 0.f;
        w.crashed = w.crashed || w.has_impact;
        w.has_impact = false;
        w.impx = w.impy = 0.f;
        w.h += w.v * sb * 0.4f * dt;   // / (LENGTH / 2); the product form saves the IEEE division
        w.v += w.acc * dt;
        w.lane = closest_lane(w.y, P.lanes);
        sincos_heading(w.h, &w.sh, &w.ch);
        if (k < V) {
            S.x[k] = w.x; S.xr[k] = (float)(w.x - xref);
            S.y[k] = w.y; S.v[k] = w.v; S.ch[k] = w.ch; S.sh[k] = w.sh;
            S.lane[k] = (unsigned char)w.lane;
            S.tl_old[k] = S.tl_new[k] = (unsigned char)w.tlane;
        }
    }
    __syncwarp();
    rank_repair(S, V, lane);

    // ---- collisions: pairs within the pre-check radius are neighbours in rank order
    int a[2];
    float xa[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        a[q] = lane + 32 * q < V ? (int)S.order[lane + 32 * q] : -1;
        xa[q] = S.xr[max(a[q], 0)];
    }
    for (int off = 1; off < V; ++off) {
        int b[2];
        bool cand[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            int s2 = lane + 32 * q + off;
            cand[q] = a[q] >= 0 && s2 < V;
            b[q] = 0;
            if (cand[q]) {
                b[q] = S.order[s2];
                cand[q] = S.xr[b[q]] - xa[q] <= 8.8f;  // >= sqrt(29) + max speed * dt
            }
        }
        if (!__any_sync(HRP_FULL, cand[0] || cand[1])) break;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            int ev = 0, ei = 0, ej = 0;
            float tx = 0.f, ty = 0.f;
            if (cand[q]) {
                ei = min(a[q], b[q]); ej = max(a[q], b[q]);
                ev = collide_pair(S, ei, ej, dt, tx, ty);
            }
            unsigned evm = __ballot_sync(HRP_FULL, ev != 0);
            while (evm) {  // rare: serialise so that the last pair in (i, j) order wins the impact
                int src = __ffs(evm) - 1;
                evm &= evm - 1;
                int bv = __shfl_sync(HRP_FULL, ev, src);
                int bi = __shfl_sync(HRP_FULL, ei, src), bj = __shfl_sync(HRP_FULL, ej, src);
                float btx = __shfl_sync(HRP_FULL, tx, src), bty = __shfl_sync(HRP_FULL, ty, src);
                if (lane == 0) {
                    if (bv & 2) {
                        int ki = 64 + bj, kj = bi;  // pairs (i, *) come after every pair (*, i)
                        if (ki > S.impkey[bi]) { S.impkey[bi] = ki; S.impx[bi] = 0.5f * btx; S.impy[bi] = 0.5f * bty; }
                        if (kj > S.impkey[bj]) { S.impkey[bj] = kj; S.impx[bj] = -0.5f * btx; S.impy[bj] = -0.5f * bty; }
                    }
                    if (bv & 1) { S.crash[bi] = 1; S.crash[bj] = 1; }
                }
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int k = lane + 32 * q;
        if (k >= V) continue;
        Veh &w = u[q];
        if (S.crash[k]) { w.crashed = true; S.crash[k] = 0; }
        if (S.impkey[k] >= 0) {
            w.has_impact = true; w.impx = S.impx[k]; w.impy = S.impy[k];
            S.impkey[k] = -1;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * HRP_WARPS_PER_CTA, HRP_STEP_CTAS_PER_SM)
hrp_step_kernel(const EnvDev P, const float *__restrict__ actions, float *__restrict__ obs,
                float *__restrict__ reward, uint8_t *__restrict__ term, uint8_t *__restrict__ trunc,
                const int32_t *__restrict__ perm, int32_t *__restrict__ row_vehicle)
{
    __shared__ WarpS smem[HRP_WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * HRP_WARPS_PER_CTA + warp;
    if (e >= P.E) return;
    WarpS &S = smem[warp];
    Veh u[2];
    double xref;
    load_env(P, S, u, e, lane, xref);
    rank_full(S, P.V, lane);

    // Launched with programmatic stream serialisation: everything above needs the simulator state only, so it may
    // overlap the tail of the policy kernel that is still producing the actions.  No hrp_pdl_release() in this
    // kernel: a following step would read the state before this one has stored it.
    hrp_pdl_wait();
    // ActionType.act on the first frame
    float a0 = actions[2 * e], a1 = actions[2 * e + 1];
    if (lane == 0) {
        if (P.ego_mode == 0) {
            // ContinuousAction.get_action: float32 lmap of the clipped np.float32 action
            a0 = clipf(a0, -1.f, 1.f); a1 = clipf(a1, -1.f, 1.f);
            u[0].acc = __fadd_rn(-5.0f, __fdiv_rn(__fmul_rn(__fsub_rn(a0, -1.0f), 10.0f), 2.0f));
            float steer = __fadd_rn(-0.78539816339744831f,
                                    __fdiv_rn(__fmul_rn(__fsub_rn(a1, -1.0f), 1.5707963267948966f), 2.0f));
            u[0].tb = 0.5f * tanf(steer);
        } else {
            // MDPVehicle.act(action) / ControlledVehicle.act(action)
            int act = (int)a0;
            Veh &w = u[0];
            if (act == 3 || act == 4) {
                int idx = speed_to_index(w.v) + (act == 3 ? 1 : -1);
                idx = max(0, min(2, idx));
                w.ts = 20.f + 5.f * idx;
                S.ts[0] = w.ts;
            } else if (act == 0 || act == 2) {
                int t = max(0, min(P.lanes - 1, w.tlane + (act == 2 ? 1 : -1)));
                if (fabsf(w.y - kLaneW * t) <= 2.f * kLaneW && w.x >= 0.0 && w.x < 10005.0) w.tlane = t;
                S.tl_old[0] = S.tl_new[0] = (unsigned char)w.tlane;
            }
        }
    }
    __syncwarp();

    for (int frame = 0; frame < P.frames; ++frame) simulate_frame(P, S, u, lane, xref);

    // HighwayEnv._reward / _is_terminated / _is_truncated
    int done = 0;
    double tnow = 0.0;
    if (lane == 0) {
        const Veh &w = u[0];
        int rl = P.ego_mode == 1 ? w.tlane : w.lane;
        float fs = w.v * w.ch;
        float scaled = (fs - P.rs_lo) / (P.rs_hi - P.rs_lo);
        bool on_road = fabsf(w.y - kLaneW * w.lane) <= 2.f && w.x >= -5.0 && w.x < 10005.0;
        float r = P.collision_reward * (w.crashed ? 1.f : 0.f) +
                  P.right_lane_reward * ((float)rl / (float)max(P.lanes - 1, 1)) +
                  P.high_speed_reward * clipf(scaled, 0.f, 1.f);
        if (P.normalize_reward)
            r = (r - P.collision_reward) / ((P.high_speed_reward + P.right_lane_reward) - P.collision_reward);
        r *= on_road ? 1.f : 0.f;
        tnow = P.time[e] + P.dtime;
        bool te = w.crashed || (P.offroad_terminal && !on_road);
        bool tr = tnow >= P.duration;
        reward[e] = r; term[e] = te; trunc[e] = tr;
        done = (te || tr) ? 1 : 0;
    }
    done = __shfl_sync(HRP_FULL, done, 0);
    uint32_t draw = P.obs_draw[e];
    if (done && P.autoreset) {
        uint32_t ep = P.episode[e] + 1;
        spawn_env(P, S, u, e, lane, ep, xref);
        if (lane == 0) { P.episode[e] = ep; tnow = 0.0; }
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (lane + 32 * q < P.V) S.h[lane + 32 * q] = u[q].h;
        __syncwarp();
    }
    if (lane == 0) { P.time[e] = tnow; P.obs_draw[e] = draw + 1; }
    observe_env(P, S, e, lane, obs, perm, row_vehicle, draw);
    store_env(P, u, e, lane);
}

__global__ void __launch_bounds__(32 * HRP_WARPS_PER_CTA)
hrp_observe_kernel(const EnvDev P, float *__restrict__ obs, const int32_t *__restrict__ perm,
                   int32_t *__restrict__ row_vehicle)
{
    __shared__ WarpS smem[HRP_WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * HRP_WARPS_PER_CTA + warp;
    if (e >= P.E) return;
    WarpS &S = smem[warp];
    Veh u[2];
    double xref;
    load_env(P, S, u, e, lane, xref);
    uint32_t draw = P.obs_draw[e];
    observe_env(P, S, e, lane, obs, perm, row_vehicle, draw);
    if (lane == 0) P.obs_draw[e] = draw + 1;
}

__global__ void __launch_bounds__(32 * HRP_WARPS_PER_CTA)
hrp_reset_kernel(const EnvDev P, const uint8_t *__restrict__ mask, float *__restrict__ obs)
{
    __shared__ WarpS smem[HRP_WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * HRP_WARPS_PER_CTA + warp;
    if (e >= P.E) return;
    if (mask && !mask[e]) return;
    WarpS &S = smem[warp];
    Veh u[2];
    double xref;
    spawn_env(P, S, u, e, lane, 0u, xref);
    uint32_t draw = P.obs_draw[e];
    if (lane == 0) { P.episode[e] = 0; P.time[e] = 0.0; P.obs_draw[e] = draw + 1; }
    if (obs) observe_env(P, S, e, lane, obs, nullptr, nullptr, draw);
    store_env(P, u, e, lane);
}

// standalone wrapper.observation() over caller-provided [B, N, F] observations
__global__ void __launch_bounds__(128)
hrp_embed_kernel(int kind, int edim, int use_euclid, int ego_idx, float max_dist,
                 const float *__restrict__ table, const float *__restrict__ obs, float *__restrict__ out,
                 long long batch, int N, int F, int Fout, const float *__restrict__ dist_override)
{
    // the [N, F] table is read straight from global memory (L1-resident, a few KB per warp),
    // so F is not limited by the simulator's HRP_MAX_FEATURES here
    __shared__ float rowd[4][HRP_MAX_OBS_ROWS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long b = (long long)blockIdx.x * 4 + warp;
    if (b >= batch) return;
    const float *src = obs + (size_t)b * N * F;
    embed_store(kind, edim, use_euclid, ego_idx, max_dist, table, src, rowd[warp], N, F, Fout, lane,
                out + (size_t)b * N * Fout, dist_override ? dist_override + (size_t)b * N : nullptr);
}

}  // namespace

int hrp_launch_step(const EnvDev &P, const float *actions, float *obs, float *reward, uint8_t *term,
                    uint8_t *trunc, const int32_t *perm, int32_t *row_vehicle, cudaStream_t s)
{
    int grid = (P.E + HRP_WARPS_PER_CTA - 1) / HRP_WARPS_PER_CTA;
    HRP_CUDA_OK(hrp_launch_pdl(hrp_step_kernel, dim3(grid), dim3(32 * HRP_WARPS_PER_CTA), 0, s, P, actions, obs, reward, term,
                               trunc, perm, row_vehicle));
    return 0;
}
int hrp_launch_observe(const EnvDev &P, float *obs, const int32_t *perm, int32_t *row_vehicle,
                       cudaStream_t s)
{
    int grid = (P.E + HRP_WARPS_PER_CTA - 1) / HRP_WARPS_PER_CTA;
    hrp_observe_kernel<<<grid, 32 * HRP_WARPS_PER_CTA, 0, s>>>(P, obs, perm, row_vehicle);
    HRP_CUDA_OK(cudaGetLastError());
    return 0;
}
int hrp_launch_reset(const EnvDev &P, const uint8_t *mask, float *obs, cudaStream_t s)
{
    int grid = (P.E + HRP_WARPS_PER_CTA - 1) / HRP_WARPS_PER_CTA;
    hrp_reset_kernel<<<grid, 32 * HRP_WARPS_PER_CTA, 0, s>>>(P, mask, obs);
    HRP_CUDA_OK(cudaGetLastError());
    return 0;
}
int hrp_launch_embed(int kind, int embed_dim, int use_euclid, int ego_idx, float max_dist,
                     const float *table, const float *obs, float *out, long long batch, int rows,
                     int cols, const float *dist_override, cudaStream_t s)
{
    int fout = kind == HRP_EMBED_DIST || kind == HRP_EMBED_RANK ? cols + embed_dim : cols;
    unsigned grid = (unsigned)((batch + 3) / 4);
    hrp_embed_kernel<<<grid, 128, 0, s>>>(kind, embed_dim, use_euclid, ego_idx, max_dist, table, obs, out,
                                          batch, rows, cols, fout, dist_override);
    HRP_CUDA_OK(cudaGetLastError());
    return 0;
}
