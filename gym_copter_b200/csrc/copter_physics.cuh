// copter_physics.cuh -- device-side arithmetic of the batched copter step: build knobs, kernel
// constants, vector-plane state access, Philox4x32-10, Eq. 6 / Eq. 12 dynamics with the
// flight-status machine, the lander reward, one reference env step, the reset state and the
// observation staging.  Included by copter_kernels.cu (kernels, launchers, C ABI).
//
// What each device function restates (paths relative to the reference root):
//   motor_forces()      gym_copter/dynamics/__init__.py:120-132, 231-247   (Eq. 6)
//   dynamics_update()   gym_copter/dynamics/__init__.py:139-197, 249-302   (Eq. 12, FSM, Euler)
//   lander_shaping()    gym_copter/envs/lander.py:48-56
//   env_substep()       gym_copter/envs/task.py:77-137 + gym_copter/envs/lander.py:58-72
//   reset state         gym_copter/envs/task.py:145-197, gym_copter/dynamics/__init__.py:210-217
//
// Precision.  T = double reproduces the numpy reference to ~1e-13.  T = float stores and
// integrates in fp32 but evaluates the motor -> thrust/torque stage in fp64: the squares of
// fp32 motor commands are exact in fp64, which removes the systematic thrust/torque bias
// that otherwise grows like t^2 (altitude) and t^4 (lateral position) and breaks the 1e-4
// budget over 1000 steps (measured: 1.5e-3 all-fp32 vs 1.8e-5 mixed).

#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/copter_b200.h"

namespace copter {


#ifndef COPTER_F32_CTAS_PER_SM
#define COPTER_F32_CTAS_PER_SM 8  // x 128 threads: 1024 resident envs per SM at <= 64 registers
#endif

#ifndef COPTER_BLOCK
#define COPTER_BLOCK 128
#endif
#ifndef COPTER_LIBM_ONLY
#define COPTER_LIBM_ONLY 0      // 1 (A/B knob): library sincosf, IEEE sqrt and division everywhere
#endif
#ifndef COPTER_PREFETCH
#define COPTER_PREFETCH 0       // (A/B knob, persistent grids only) request the NEXT tile's loads before this tile's arithmetic
#endif
#ifndef COPTER_PERSISTENT
#define COPTER_PERSISTENT 0     // 1 (A/B knob): one resident wave of CTAs walking the tiles with a grid stride
#endif
#ifndef COPTER_K1_SPECIALIZE
#define COPTER_K1_SPECIALIZE 1   // dedicated code path for k_substeps == 1 (the HBM-bound case)
#endif
#ifndef COPTER_K_UNROLL
#define COPTER_K_UNROLL 1         // unroll factor of the substep loop (A/B knob)
#endif
#ifndef COPTER_FAST_SUBSTEP
#define COPTER_FAST_SUBSTEP 1     // 0 (A/B knob): every substep of a K-fused launch takes the general env_advance
#endif
#ifndef COPTER_CALM_STREAK
#define COPTER_CALM_STREAK 1     // 0 (A/B knob): every straight-line substep re-derives the ending flags and the hot test
#endif
#ifndef COPTER_STREAMING
#define COPTER_STREAMING 0      // 1: evict-first (ld/st .cs) hints on the state planes
#endif

constexpr int kBlock = COPTER_BLOCK;
constexpr int kWarpsPerBlock = kBlock / 32;

enum { ST_CRASHED = 0, ST_LANDED = 1, ST_LEVELING = 2, ST_AIRBORNE = 3 };
enum { CAUSE_LANDED = 1, CAUSE_BONUS = 2, CAUSE_OOB = 4, CAUSE_ANGLE = 8, CAUSE_CRASHED = 16, CAUSE_TIMEOUT = 32 };

// ------------------------------------------------------------------------------------------
// compile-time description of the env variants
// ------------------------------------------------------------------------------------------
template <int VARIANT> struct Variant;
template <> struct Variant<COPTER_LANDER3D> { static constexpr int O = 10, A = 4, first = 0; static constexpr bool lander = true; };
template <> struct Variant<COPTER_LANDER2D> { static constexpr int O = 6,  A = 2, first = 2; static constexpr bool lander = true; };
template <> struct Variant<COPTER_LANDER1D> { static constexpr int O = 2,  A = 1, first = 4; static constexpr bool lander = true; };
template <> struct Variant<COPTER_HOVER3D>  { static constexpr int O = 12, A = 4, first = 0; static constexpr bool lander = false; };
template <> struct Variant<COPTER_HOVER2D>  { static constexpr int O = 6,  A = 2, first = 2; static constexpr bool lander = false; };
template <> struct Variant<COPTER_HOVER1D>  { static constexpr int O = 2,  A = 1, first = 4; static constexpr bool lander = false; };

// ------------------------------------------------------------------------------------------
// kernel-side constants, derived once on the host from CopterParams
// ------------------------------------------------------------------------------------------
template <typename T>
struct KParams {
    double kT, kR, kP, kY;        // B w^2/M, L B w^2/Ix, L B w^2/Iy, D w^2/Iz  (w = maxrpm*pi/30)
    double kOm;                   // w when the gyroscopic coupling is live (COPTER_MODEL_GYRO), else 0
    double force_scale, force_off; // u32 -> U(-F,F): u * 2F/2^32 - F
    T G, dt, gphi, gthe, gpsi;    // (Iy-Iz)/Ix, (Iz-Ix)/Iy, (Ix-Iy)/Iz
    T lvx, lvy, lang, invM;
    T jx, jy;                     // Jr/Ix, Jr/Iy
    T oob_penalty, max_angle, bounds, z0, target_radius;
    T calm_angle;            // min(max_angle, polynomial sin/cos range): below it a step neither ends over-angle nor leaves the fast path
    T yaw_pf, xyz_pf, dz_max, dz_penalty, bonus;
    int max_steps;
    int status0;                  // status right after reset (dynamics/__init__.py:215-217)
};

template <typename T>
KParams<T> make_kparams(const CopterParams& p) {
    KParams<T> k;
    const double w = p.maxrpm * M_PI / 30.0;
    // thrust per unit w^2 and the roll/pitch torque arm: live model B and L (dynamics/__init__.py:127-129),
    // lift model 0.5 rho S C_L (L/2)^2 and 1 (attic/mars/dynamics/__init__.py:101,146-158)
    const bool lift = (p.dynamics_model & COPTER_MODEL_LIFT) != 0;
    const double b = lift ? 0.5 * p.rho * (0.05 * p.L * 4) * p.lift_coefficient * (p.L / 2) * (p.L / 2) : p.B;
    const double arm = lift ? 1.0 : p.L;
    k.kT = b * w * w / p.M;
    k.kR = arm * b * w * w / p.Ix;
    k.kP = arm * b * w * w / p.Iy;
    k.kY = p.D * w * w / p.Iz;
    k.kOm = (p.dynamics_model & COPTER_MODEL_GYRO) ? w : 0.0;
    k.jx = (T)(p.Jr / p.Ix); k.jy = (T)(p.Jr / p.Iy);
    k.force_scale = 2.0 * p.initial_random_force / 4294967296.0;
    k.force_off = -p.initial_random_force;
    k.G = (T)p.G;
    k.dt = (T)((T)1 / (T)p.fps);
    k.gphi = (T)((p.Iy - p.Iz) / p.Ix);
    k.gthe = (T)((p.Iz - p.Ix) / p.Iy);
    k.gpsi = (T)((p.Ix - p.Iy) / p.Iz);
    k.lvx = (T)p.landing_vel_x; k.lvy = (T)p.landing_vel_y; k.lang = (T)p.landing_angle;
    k.invM = (T)(1.0 / p.M);
    k.oob_penalty = (T)p.out_of_bounds_penalty;
    k.max_angle = (T)(p.max_angle_deg * M_PI / 180.0);
    k.calm_angle = (sizeof(T) == 4 && k.max_angle > (T)0.78539816f) ? (T)0.78539816f : k.max_angle;
    k.bounds = (T)p.bounds;
    k.z0 = (T)(-p.initial_altitude);
    k.target_radius = (T)p.target_radius;
    k.yaw_pf = (T)p.yaw_penalty_factor; k.xyz_pf = (T)p.xyz_penalty_factor;
    k.dz_max = (T)p.dz_max; k.dz_penalty = (T)p.dz_penalty; k.bonus = (T)p.inside_radius_bonus;
    k.max_steps = p.max_steps;
    k.status0 = (-p.initial_altitude < 0) ? ST_AIRBORNE : ST_LANDED;
    return k;
}

// ------------------------------------------------------------------------------------------
// small typed helpers
// ------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float>  { using type = float4;  static constexpr int V = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int V = 2; };

// fp32 sin/cos.  |a| <= pi/4 needs no range reduction: evaluate the same degree-7 / degree-8
// minimax polynomials the accurate sincosf uses on its reduced interval (max rel. error
// 7e-8 / 9e-8 over the interval) and skip its quadrant logic; anything larger takes the
// library's accurate path (roll/pitch beyond pi/4 end the episode, task.py:116, so in
// practice only a large yaw angle ever does).  Never the SFU approximations (__sinf/__cosf).
__device__ __forceinline__ void sincos_t(float a, float* s, float* c) {
    if (!COPTER_LIBM_ONLY && fabsf(a) <= 0.78539816f) {
        const float z = a * a;
        float ps = fmaf(z, -1.95152959e-4f, 8.33216087e-3f);
        ps = fmaf(ps, z, -1.66666546e-1f);
        *s = fmaf(a * z, ps, a);
        float pc = fmaf(z, 2.44331571e-5f, -1.38873163e-3f);
        pc = fmaf(pc, z, 4.16666456e-2f);
        pc = fmaf(pc, z, -0.5f);
        *c = fmaf(pc, z, 1.0f);
    } else {
        sincosf(a, s, c);
    }
}
__device__ __forceinline__ void sincos_t(double a, double* s, double* c) { sincos(a, s, c); }

// sin/cos of roll, pitch and yaw together: ONE range test for the three angles on the fp32
// path (all three are below pi/4 in every step that matters), then three polynomial pairs.
__device__ __forceinline__ void sincos_poly(float a, float& s, float& c) {
    const float z = a * a;
    float ps = fmaf(z, -1.95152959e-4f, 8.33216087e-3f);
    ps = fmaf(ps, z, -1.66666546e-1f);
    s = fmaf(a * z, ps, a);
    float pc = fmaf(z, 2.44331571e-5f, -1.38873163e-3f);
    pc = fmaf(pc, z, 4.16666456e-2f);
    pc = fmaf(pc, z, -0.5f);
    c = fmaf(pc, z, 1.0f);
}
__device__ __forceinline__ void sincos3_t(float a, float b, float g, float& sa, float& ca, float& sb, float& cb, float& sg, float& cg) {
    if (!COPTER_LIBM_ONLY && fmaxf(fmaxf(fabsf(a), fabsf(b)), fabsf(g)) <= 0.78539816f) {
        sincos_poly(a, sa, ca); sincos_poly(b, sb, cb); sincos_poly(g, sg, cg);
    } else {
        sincosf(a, &sa, &ca); sincosf(b, &sb, &cb); sincosf(g, &sg, &cg);
    }
}
__device__ __forceinline__ void sincos3_t(double a, double b, double g, double& sa, double& ca, double& sb, double& cb, double& sg, double& cg) {
    sincos(a, &sa, &ca); sincos(b, &sb, &cb); sincos(g, &sg, &cg);
}
__device__ __forceinline__ float  sqrt_t(float a)  { return sqrtf(a); }
__device__ __forceinline__ double sqrt_t(double a) { return sqrt(a); }
// Reward-only helpers (never used for the state): on the fp32 path sqrt and the quotient of
// shaping_delta go through MUFU.RSQ / MUFU.RCP (<= 2 ulp), a few 1e-7 of the reward against a
// 1e-4 budget; the fp64 path keeps IEEE sqrt and division.
__device__ __forceinline__ float  reward_sqrt(float a)  { return COPTER_LIBM_ONLY ? sqrtf(a) : (a > 0.0f ? a * rsqrtf(a) : 0.0f); }
__device__ __forceinline__ double reward_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float  reward_div(float n, float d)   { return COPTER_LIBM_ONLY ? n / d : __fdividef(n, d); }
__device__ __forceinline__ double reward_div(double n, double d) { return n / d; }
__device__ __forceinline__ float  abs_t(float a)  { return fabsf(a); }
__device__ __forceinline__ double abs_t(double a) { return fabs(a); }

template <typename T>
__device__ __forceinline__ void load_state(const T* __restrict__ state, int64_t stride, int64_t i, T (&s)[12]) {
    using V4 = typename Vec<T>::type;
    constexpr int V = Vec<T>::V;
    const V4* planes = reinterpret_cast<const V4*>(state);
#pragma unroll
    for (int pl = 0; pl < 12 / V; ++pl) {
        V4 v = COPTER_STREAMING ? __ldcs(&planes[(int64_t)pl * stride + i]) : planes[(int64_t)pl * stride + i];
        const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int j = 0; j < V; ++j) s[pl * V + j] = e[j];
    }
}

template <typename T>
__device__ __forceinline__ void store_state(T* __restrict__ state, int64_t stride, int64_t i, const T (&s)[12]) {
    using V4 = typename Vec<T>::type;
    constexpr int V = Vec<T>::V;
    V4* planes = reinterpret_cast<V4*>(state);
#pragma unroll
    for (int pl = 0; pl < 12 / V; ++pl) {
        V4 v;
        T* e = reinterpret_cast<T*>(&v);
#pragma unroll
        for (int j = 0; j < V; ++j) e[j] = s[pl * V + j];
        if (COPTER_STREAMING) __stcs(&planes[(int64_t)pl * stride + i], v); else planes[(int64_t)pl * stride + i] = v;
    }
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11), counter-based: no per-env generator state in HBM
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// Reset force for (global env id, episode): exact in fp64, ONE rounding to T.
template <typename T>
__device__ __forceinline__ void reset_force(const KParams<T>& kp, uint64_t seed, uint64_t env, uint32_t episode, T (&f)[3]) {
    uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), episode, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int j = 0; j < 3; ++j) f[j] = (T)fma((double)c[j], kp.force_scale, kp.force_off);
}

// ------------------------------------------------------------------------------------------
// dynamics
// ------------------------------------------------------------------------------------------
template <typename T> struct Forces { T bz, u2, u3, u4, om; };   // -U1/M, U2/Ix, U3/Iy, U4/Iz, Omega

// dynamics/__init__.py:120-132.  Always evaluated in fp64 (see header comment).
template <typename T>
__device__ __forceinline__ Forces<T> motor_forces(const KParams<T>& kp, T m0, T m1, T m2, T m3) {
    const double q0 = (double)m0 * (double)m0, q1 = (double)m1 * (double)m1;
    const double q2 = (double)m2 * (double)m2, q3 = (double)m3 * (double)m3;
    const double s01 = q0 + q1, s23 = q2 + q3;
    Forces<T> f;
    f.bz = (T)(-kp.kT * (s01 + s23));
    f.u2 = (T)(kp.kR * ((q1 + q2) - (q0 + q3)));      // roll right  (:231-235)
    f.u3 = (T)(kp.kP * ((q1 + q3) - (q0 + q2)));      // pitch forward (:237-241)
    f.u4 = (T)(kp.kY * (s01 - s23));                  // yaw cw (:243-247)
    // Omega: zero in the live model (:135); u4 of the UNSQUARED speeds in attic/mars (:143)
    f.om = (T)(kp.kOm * (((double)m0 + (double)m1) - ((double)m2 + (double)m3)));
    return f;
}

// dynamics/__init__.py:139-197 for one env.  NP = number of perturbed rate components the
// caller supplies (3 on the env path: x,y,z only, envs/task.py:179-184; 6 for the Dynamics
// facade).  DIRECT enables the LANDED -> AIRBORNE take-off transition, unreachable through
// _Task.step (task.py:86-94).  Returns true when the call ran to the end of setMotors
// (perturbation cleared, ticks += 1), false on the ground-contact early return (:177).
// `na` / `nc` receive the shaping numerators sum_j inc_j (2 s_j + inc_j) over (x,dx,y,dy,z,dz)
// and (psi,dpsi), inc_j = dt*ds_j being the Euler increment BEFORE it is rounded into the
// state (zero when the state was not integrated) -- see shaping_delta.
// The hot case (AIRBORNE, not touching the ground) is tested first and is straight-line code.
// The AIRBORNE, not-touching-the-ground case of setMotors (:180-197): Eq. 12 and one forward Euler
// step from the sines / cosines of the current angles.  PERT: the reset perturbation `p` is added
// (twice, :263-287 and :183) to the first NP rate derivatives.  Shared by dynamics_update and the
// straight-line substep of the K-fused loop (airborne_substep), so both produce the same bits.
template <typename T, int NP, bool PERT>
__device__ __forceinline__ void airborne_integrate(const KParams<T>& kp, T (&s)[12], const Forces<T>& f, const T (&p)[NP],
                                                   T sph, T cph, T sth, T cth, T sps, T cps, T& na, T& nc) {
    // third column of the body->inertial rotation times the body-Z thrust (:292-302)
    const T ax = f.bz * (sph * sps + cph * cps * sth);
    const T ay = f.bz * (cph * sps * sth - cps * sph);
    const T netz = f.bz * (cph * cth) + kp.G;                     // :143
    const T dphi = s[7], dthe = s[9], dpsi = s[11];
    // Eq. 12 (:257-290) with Omega = 0 (:135); the perturbation is added twice (:263-287, :183)
    T d1 = ax, d3 = ay, d5 = netz;
    T d7 = dpsi * dthe * kp.gphi - kp.jx * dthe * f.om + f.u2;
    T d9 = -(dpsi * dphi * kp.gthe + kp.jy * dphi * f.om + f.u3);
    T d11 = dthe * dphi * kp.gpsi + f.u4;
    if constexpr (PERT) {
        d1 += (T)2 * p[0]; d3 += (T)2 * p[1]; d5 += (T)2 * p[2];
        if constexpr (NP == 6) { d7 += (T)2 * p[3]; d9 += (T)2 * p[4]; d11 += (T)2 * p[5]; }
    }
    // forward Euler, every derivative from the old state (:187)
    const T dt = kp.dt;
    const T i0 = dt * s[1], i1 = dt * d1, i2 = dt * s[3], i3 = dt * d3, i4 = dt * s[5], i5 = dt * d5;
    const T i10 = dt * dpsi, i11 = dt * d11;
    na = i0 * ((T)2 * s[0] + i0) + i1 * ((T)2 * s[1] + i1) + i2 * ((T)2 * s[2] + i2)
       + i3 * ((T)2 * s[3] + i3) + i4 * ((T)2 * s[4] + i4) + i5 * ((T)2 * s[5] + i5);
    nc = i10 * ((T)2 * s[10] + i10) + i11 * ((T)2 * s[11] + i11);
    s[0] += dt * s[1];  s[1] += dt * d1;
    s[2] += dt * s[3];  s[3] += dt * d3;
    s[4] += dt * s[5];  s[5] += dt * d5;
    s[6] += dt * dphi;  s[7] += dt * d7;
    s[8] += dt * dthe;  s[9] += dt * d9;
    s[10] += dt * dpsi; s[11] += dt * d11;
}

template <typename T, int NP, bool DIRECT>
__device__ __forceinline__ bool dynamics_update(const KParams<T>& kp, T (&s)[12], int& st,
                                                const Forces<T>& f, const T (&p)[NP], T& na, T& nc) {
    na = (T)0; nc = (T)0;
    T sph, cph, sth, cth, sps, cps;
    sincos3_t(s[6], s[8], s[10], sph, cph, sth, cth, sps, cps);

    if (DIRECT && st == ST_LANDED) {                               // :147-149
        const T netz = f.bz * (cph * cth) + kp.G;                  // :143
        if (netz < (T)0) st = ST_AIRBORNE;
    }

    const bool touch = s[4] > (T)0 && s[5] > (T)0;                 // :162 (pre-step state)
    if (st == ST_AIRBORNE && !touch) {                             // :159, :180-187
        airborne_integrate<T, NP, true>(kp, s, f, p, sph, cph, sth, cth, sps, cps, na, nc);
        return true;
    }
    if (st == ST_LEVELING) {                                       // :152-156
        s[6] = (T)0; s[8] = (T)0; st = ST_LANDED;
        return true;
    }
    if (st == ST_AIRBORNE) {                                       // touched the ground (:162-177)
        // :165-171 -- "velx" is dy, "vely" is dz, only phi is angle-tested (sic)
        st = (s[5] > kp.lvy || abs_t(s[3]) > kp.lvx || abs_t(s[6]) > kp.lang) ? ST_CRASHED : ST_LEVELING;
        return false;                                              // :177
    }
    return true;
}

// envs/lander.py:48-56, kept as its three ingredients: shaping = -(xyz_pf*ra + yaw_pf*rc) - pen
template <typename T> struct Shaping { T ra, rc, pen; };

template <typename T>
__device__ __forceinline__ Shaping<T> lander_shaping(const KParams<T>& kp, const T (&s)[12]) {
    const T spos = ((((s[0] * s[0] + s[1] * s[1]) + s[2] * s[2]) + s[3] * s[3]) + s[4] * s[4]) + s[5] * s[5];
    const T spsi = s[10] * s[10] + s[11] * s[11];
    Shaping<T> sh;
    sh.ra = reward_sqrt(spos);
    sh.rc = reward_sqrt(spsi);
    sh.pen = abs_t(s[5]) > kp.dz_max ? kp.dz_penalty : (T)0;
    return sh;
}

// reward = shaping(post) - shaping(pre) (envs/lander.py:58-62), evaluated without the
// cancellation of two O(250..1e4) numbers:  sqrt(a1) - sqrt(a0) = (a1 - a0) / (sqrt(a1) + sqrt(a0))
// with a1 - a0 = sum_j inc_j (2 pre_j + inc_j) (`na`, `nc` from dynamics_update), where
// inc_j = dt*ds_j is the Euler increment BEFORE it is rounded into the stored state.  In fp32
// this keeps the reward error proportional to |reward| (1e-5 measured) instead of
// |shaping| * 2^-24 (literal subtraction, up to 1e-3) or ulp(state)/increment (differences of
// stored states, 3e-4 at |v| ~ 270 m/s); in fp64 it agrees with the reference's literal
// subtraction to ~1e-13.
template <typename T>
__device__ __forceinline__ T shaping_delta(const KParams<T>& kp, const Shaping<T>& pre, T na, T nc,
                                           const Shaping<T>& post) {
    const T da = post.ra + pre.ra, dc = post.rc + pre.rc;
    const T ga = da > (T)0 ? reward_div(na, da) : (T)0;
    const T gc = dc > (T)0 ? reward_div(nc, dc) : (T)0;
    return -(kp.xyz_pf * ga + kp.yaw_pf * gc) - (post.pen - pre.pen);
}

// One reference _Task.step (envs/task.py:77-137) for one env held in registers, WITHOUT the
// reward: advances the dynamics, the status machine and the step counter and reports whether
// the episode ended and why.  `na` / `nc` are the shaping numerators of this step (see
// dynamics_update); the reward modifiers are encoded in `cause` (BONUS: + bonus, lander.py:69-72;
// OOB: - penalty, task.py:111-113; ANGLE: reward := - penalty, task.py:116-118; the two are
// exclusive because the reference tests them with if / elif).
template <typename T, int VARIANT>
__device__ __forceinline__ void env_advance(const KParams<T>& kp, T (&s)[12], int& st, int& steps,
                                            const Forces<T>& f, const T (&pert)[3], T& na, T& nc,
                                            bool& done, int& cause) {
    const int st0 = st;                                            // :81 stale status
    na = (T)0; nc = (T)0;
    if (st0 != ST_LANDED)                                          // :86-94
        dynamics_update<T, 3, false>(kp, s, st, f, pert, na, nc);
    cause = 0;
    done = false;
    if (Variant<VARIANT>::lander && st0 == ST_LANDED) {            // lander.py:64-72
        done = true; cause |= CAUSE_LANDED;
        if (sqrt_t(s[0] * s[0] + s[2] * s[2]) < kp.target_radius) cause |= CAUSE_BONUS;
    }
    if (abs_t(s[0]) >= kp.bounds || abs_t(s[2]) >= kp.bounds) {    // task.py:111
        done = true; cause |= CAUSE_OOB;
    } else if (abs_t(s[6]) >= kp.max_angle || abs_t(s[8]) >= kp.max_angle) {   // :116
        done = true; cause |= CAUSE_ANGLE;
    } else if (st0 == ST_CRASHED) {                                // :121
        done = true;
    }
    if (st0 == ST_CRASHED) cause |= CAUSE_CRASHED;
    if (steps == kp.max_steps) { done = true; cause |= CAUSE_TIMEOUT; }          // :128
    steps = min(steps + 1, 2047);                                  // :130 (11-bit field)
    if (!done) cause = 0;
}

// The common case of env_advance as straight-line code, for the K-fused loops: an AIRBORNE env that
// is not touching the ground, is past the first step of its episode (no reset perturbation left)
// and -- fp32 -- has all three angles inside the polynomial range of sincos_poly.  airborne_hot()
// is that precondition; under it airborne_substep() gives exactly what env_advance gives.
template <typename T>
__device__ __forceinline__ bool airborne_hot(const T (&s)[12], int st, int steps) {
    // (bitwise on purpose: one straight run of compares, no short-circuit branches)
    int hot = (int)(st == ST_AIRBORNE) & (int)(steps != 1) & (int)!(s[4] > (T)0 && s[5] > (T)0);
    if constexpr (sizeof(T) == 4) hot &= (int)(!COPTER_LIBM_ONLY && fmaxf(fmaxf(fabsf(s[6]), fabsf(s[8])), fabsf(s[10])) <= 0.78539816f);
    return hot != 0;
}

// Returns the step's ending flags: bit 0 out of bounds, bit 1 over-angle (exclusive: the reference
// tests them with if / elif, task.py:111-118), bit 2 the env's own step limit; 0 = the episode goes
// on.  airborne_cause() turns them into the CAUSE_* bits env_advance reports.
enum { END_OOB = 1, END_ANGLE = 2, END_TIMEOUT = 4 };
// the arithmetic of one such step (no flags, no step counter)
template <typename T>
__device__ __forceinline__ void airborne_arith(const KParams<T>& kp, T (&s)[12], const Forces<T>& f, T& na, T& nc) {
    T sph, cph, sth, cth, sps, cps;
    if constexpr (sizeof(T) == 4) { sincos_poly(s[6], sph, cph); sincos_poly(s[8], sth, cth); sincos_poly(s[10], sps, cps); }
    else sincos3_t(s[6], s[8], s[10], sph, cph, sth, cth, sps, cps);
    const T none[3] = {(T)0, (T)0, (T)0};
    airborne_integrate<T, 3, false>(kp, s, f, none, sph, cph, sth, cth, sps, cps, na, nc);
}
// task.py:111-130 with the stale status AIRBORNE, on the state after the step
template <typename T>
__device__ __forceinline__ int airborne_flags(const KParams<T>& kp, const T (&s)[12], bool timeout) {
    const bool oob = abs_t(s[0]) >= kp.bounds || abs_t(s[2]) >= kp.bounds;
    const bool ang = !oob && (abs_t(s[6]) >= kp.max_angle || abs_t(s[8]) >= kp.max_angle);
    return (oob ? END_OOB : 0) | (ang ? END_ANGLE : 0) | (timeout ? END_TIMEOUT : 0);
}
template <typename T, int VARIANT>
__device__ __forceinline__ int airborne_substep(const KParams<T>& kp, T (&s)[12], int& steps, const Forces<T>& f,
                                                T& na, T& nc) {
    airborne_arith<T>(kp, s, f, na, nc);
    const bool timeout = steps == kp.max_steps;
    steps = min(steps + 1, 2047);
    return airborne_flags<T>(kp, s, timeout);
}
// "Calm" after a straight-line step: the step ended nothing (in bounds, under the angle limit, not
// the step limit) AND the env is still in airborne_hot()'s common case, so the next substep can be
// straight-line too.  Deliberately a little stricter than the two conditions it implies (psi is held
// to the angle limit as well, `<` instead of `<=` at the polynomial range): a lane that is not calm
// just goes back to the exact tests.  Five compares instead of the flags + hot + two warp votes.
template <typename T>
__device__ __forceinline__ bool airborne_calm(const KParams<T>& kp, const T (&s)[12], bool timeout) {
    const T m_xy = fmax(abs_t(s[0]), abs_t(s[2]));
    const T m_ang = fmax(fmax(abs_t(s[6]), abs_t(s[8])), abs_t(s[10]));
    return (int)!timeout & (int)(m_xy < kp.bounds) & (int)(m_ang < kp.calm_angle) & (int)!(s[4] > (T)0 && s[5] > (T)0);
}
__device__ __forceinline__ int airborne_cause(int end) {
    return ((end & END_OOB) ? CAUSE_OOB : 0) | ((end & END_ANGLE) ? CAUSE_ANGLE : 0) | ((end & END_TIMEOUT) ? CAUSE_TIMEOUT : 0);
}

// The reward modifiers of task.py:111-118 and lander.py:69-72 applied to a base reward.
template <typename T>
__device__ __forceinline__ T apply_reward_modifiers(const KParams<T>& kp, T r, int cause) {
    if (cause & CAUSE_BONUS) r += kp.bonus;
    if (cause & CAUSE_OOB) r -= kp.oob_penalty;
    else if (cause & CAUSE_ANGLE) r = -kp.oob_penalty;
    return r;
}

// env_advance + the step's reward.  `pre_sh` is shaping(pre-step state) == prev_shaping (the
// priming step of _reset sets it to shaping(s0) and every later step stores the post-step
// value, task.py:197, lander.py:62), so it never has to live in HBM; on return it holds
// shaping(post).
template <typename T, int VARIANT>
__device__ __forceinline__ void env_substep(const KParams<T>& kp, T (&s)[12], int& st, int& steps,
                                            const Forces<T>& f, const T (&pert)[3], Shaping<T>& pre_sh,
                                            T& reward, bool& done, int& cause) {
    T na, nc;
    env_advance<T, VARIANT>(kp, s, st, steps, f, pert, na, nc, done, cause);
    if (Variant<VARIANT>::lander) {
        const Shaping<T> sh = lander_shaping<T>(kp, s);            // lander.py:48-56
        reward = shaping_delta<T>(kp, pre_sh, na, nc, sh);         // :58-62
        pre_sh = sh;
    } else {
        reward = (T)1;                                             // attic hover.py:18-21
    }
    reward = apply_reward_modifiers<T>(kp, reward, cause);
}

// ------------------------------------------------------------------------------------------
// Telescoped reward of a run of consecutive steps of ONE episode.  sum_k (shaping_k -
// shaping_{k-1}) = shaping_end - shaping_start, so a fused loop only accumulates the shaping
// numerators (two adds per step) and the square roots / quotients are evaluated once, by
// segment_reward(), when the run ends (episode finished, or last step of the launch).
// An over-angle ending REPLACES its own step reward by the penalty (task.py:116-118): that
// step's numerators are then left out of the sums and the run ends at the state before it,
// recovered from a_prev = a_now - na, c_prev = c_now - nc and the previous dz.
// Hover variants: +1 per step (attic hover.py:18-21) with the same modifiers.
// ------------------------------------------------------------------------------------------
template <typename T> struct RewardRun { Shaping<T> start; T na, nc; int steps; };

template <typename T, int VARIANT>
__device__ __forceinline__ void run_begin(const KParams<T>& kp, RewardRun<T>& run, const T (&s)[12]) {
    if (Variant<VARIANT>::lander) run.start = lander_shaping<T>(kp, s);
    run.na = (T)0; run.nc = (T)0; run.steps = 0;
}

// one executed step: `cause` is the step's ending cause (0 if the episode goes on)
template <typename T>
__device__ __forceinline__ void run_step(RewardRun<T>& run, T na, T nc, int cause) {
    ++run.steps;
    if (!(cause & CAUSE_ANGLE)) { run.na += na; run.nc += nc; }
}

// reward of the run; `s` is the state after its last step (before any auto-reset), `cause` /
// `na` / `nc` / `dz_prev` belong to that last step
template <typename T, int VARIANT>
__device__ __forceinline__ T run_reward(const KParams<T>& kp, const RewardRun<T>& run, const T (&s)[12],
                                        int cause, T na, T nc, T dz_prev) {
    const bool replaced = (cause & CAUSE_ANGLE) != 0;
    if (Variant<VARIANT>::lander) {
        Shaping<T> end = lander_shaping<T>(kp, s);
        if (replaced) {
            end.ra = reward_sqrt(fmax(end.ra * end.ra - na, (T)0));
            end.rc = reward_sqrt(fmax(end.rc * end.rc - nc, (T)0));
            end.pen = abs_t(dz_prev) > kp.dz_max ? kp.dz_penalty : (T)0;
        }
        const T total = shaping_delta<T>(kp, run.start, run.na, run.nc, end);
        return replaced ? total - kp.oob_penalty : apply_reward_modifiers<T>(kp, total, cause);
    }
    const T total = (T)run.steps;
    return replaced ? total - (T)1 - kp.oob_penalty : apply_reward_modifiers<T>(kp, total, cause);
}

template <typename T>
__device__ __forceinline__ void reset_state(const KParams<T>& kp, T (&s)[12], int& st, int& steps) {
    // envs/task.py:149,164-171,191,197 and dynamics/__init__.py:215-217
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = (T)0;
    s[4] = kp.z0;
    st = kp.status0;
    steps = 1;
}

// Row-major float32 observation of a warp's 32 envs, staged through shared memory so the
// global stores are contiguous 8-byte-per-lane warp stores.  `tile` is this warp's
// 32*O-float region; `row0` the first env of the warp; `rows` how many of its envs exist.
template <int VARIANT, typename T>
__device__ __forceinline__ void write_obs_rows(float* __restrict__ obs, float* tile, int lane,
                                               int64_t row0, int rows, const T (&s)[12]) {
    constexpr int O = Variant<VARIANT>::O, first = Variant<VARIANT>::first, H = O / 2;
    float2* t2 = reinterpret_cast<float2*>(tile);
#pragma unroll
    for (int j = 0; j < H; ++j)
        t2[lane * H + j] = make_float2((float)s[first + 2 * j], (float)s[first + 2 * j + 1]);
    __syncwarp();
    float2* out = reinterpret_cast<float2*>(obs + row0 * O);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const int e = j * 32 + lane;
        if (e < rows * H) out[e] = t2[e];
    }
    __syncwarp();
}

}  // namespace copter
