// copter_physics.cuh -- device-side glue around the shared arithmetic (copter_core.h): build knobs,
// vector-plane state access, the packed two-env lane type (fma.rn.f32x2), the straight-line /
// calm-streak substep pieces of the K-fused loops, and the observation staging.
// Included by copter_kernels.cu (kernels, launchers, C ABI).
//
// The arithmetic itself -- Eq. 6 / Eq. 12 dynamics with the flight-status machine, the lander
// reward, one reference env step, the reset state, Philox -- is in copter_core.h, which also
// compiles on the host: oracle/copter_host.cpp is the fp32 / fp64 CPU restatement the kernels are
// compared with (fp32: state, status, step counters, episode indices and done flags bit for bit).

#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "copter_core.h"

namespace copter {

#ifndef COPTER_F32_CTAS_PER_SM
#define COPTER_F32_CTAS_PER_SM 8  // x 128 threads: 1024 resident envs per SM at <= 64 registers (the drawn-command rollout kernels)
#endif
// The step kernel itself: 10 CTAs per SM (<= 48 registers, which the fp32 step kernels reach without spills since the
// two-FMA shaping numerators; at 8 the compiler took 56 and the hardware placed 9).  HBM-bound launches do not care
// (K = 1: 0.4135 vs 0.4134 ms); the issue-bound ones gain a little from the extra warps (K = 8 0.834 -> 0.818, K = 16
// 1.390 -> 1.359 ms, profiles/r2_ab_k_loop5.txt; 9: 1.373, 7: 1.404; 11+ spill).
#ifndef COPTER_STEP_CTAS_PER_SM
#define COPTER_STEP_CTAS_PER_SM 10
#endif

#ifndef COPTER_BLOCK
#define COPTER_BLOCK 128
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
#ifndef COPTER_FRESH_FAST
#define COPTER_FRESH_FAST 0      // 1 (A/B knob): a warp with a fresh episode in it takes a straight-line step WITH the reset perturbation
                                 // at substep 0 instead of the general step.  Measured slower (K = 16: 1.484 vs 1.443 ms): off.
#endif
#ifndef COPTER_GROUND_SPLIT
#define COPTER_GROUND_SPLIT 0    // 1 (A/B knob): while a lane on the ground is all that keeps a warp from the straight-line step, the airborne
                                 // lanes take it and that lane runs the status machine beside them.  Measured slower (K = 16: 1.511 vs
                                 // 1.444 ms, profiles/r2_ab_k_loop.txt): the general step is only ~35 % dearer than the straight-line one,
                                 // two divergent passes cost more.  Off.
#endif
#ifndef COPTER_GROUND_FF
#define COPTER_GROUND_FF 0       // 1 (A/B knob): a lane that has reached the ground takes its remaining status-machine steps at once, alone, and
                                 // leaves the loop, instead of one per substep with the whole warp in the general step (5.7 of the 16
                                 // substeps of a K = 16 launch on a desynchronised batch).  Bit-identical, measured SLOWER (K = 16: 1.533 vs
                                 // 1.391 ms, K = 2: 0.485 vs 0.453, profiles/r2_ab_k_loop2.txt): a lone lane's env_advance costs the
                                 // warp almost what a general step of all 32 lanes does.  2: the same with the status machine written
                                 // out for a grounded vehicle (ground_advance, ~25 instructions per step): 1.462 vs 1.392 ms, K = 2 0.479
                                 // vs 0.466 (profiles/r2_ab_k_loop4.txt) -- still slower, and already at K = 2 where hardly any vehicle
                                 // lands: the extra block costs the loop more than the general steps it saves.  Off.
#endif
#ifndef COPTER_CALM_STREAK
#define COPTER_CALM_STREAK 1     // 0 (A/B knob): every straight-line substep re-derives the ending flags and the hot test
#endif
#ifndef COPTER_STREAMING
#define COPTER_STREAMING 0      // 1: evict-first (ld/st .cs) hints on the state planes
#endif
#ifndef COPTER_PAIR_MIN_K
#define COPTER_PAIR_MIN_K 0     // k_substeps >= this (fp32) use the two-envs-per-thread packed kernel (0: never; run-time override:
                                // COPTER_B200_PAIR_MIN_K in the environment).  Measured slower than the scalar loop: off.
#endif

constexpr int kBlock = COPTER_BLOCK;
constexpr int kWarpsPerBlock = kBlock / 32;

// ------------------------------------------------------------------------------------------
// small typed helpers
// ------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float>  { using type = float4;  static constexpr int V = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int V = 2; };

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

// meta word(s) <-> (status, steps, episode); `hi` is the wide-counter array or null (include/copter_b200.h)
__device__ __forceinline__ void decode_meta(uint32_t m, const uint32_t* hi, int64_t i, int& st, int& steps, uint32_t& episode) {
    st = (int)(m & 3u);
    if (hi) { steps = (int)(m >> 2); episode = hi[i]; }
    else    { steps = (int)((m >> 2) & 2047u); episode = m >> 13; }
}
__device__ __forceinline__ void store_meta(uint32_t* meta, uint32_t* hi, int64_t i, int st, int steps, uint32_t episode) {
    if (hi) { meta[i] = (uint32_t)st | ((uint32_t)steps << 2); hi[i] = episode; }
    else    meta[i] = (uint32_t)st | ((uint32_t)steps << 2) | (episode << 13);
}

// ------------------------------------------------------------------------------------------
// F2: a lane of TWO envs, one packed 64-bit register pair per quantity, every operation one FFMA2
// (fma.rn.f32x2: two independent round-to-nearest IEEE operations in one issue slot), so
// airborne_integrate<F2> produces exactly the bits of two airborne_integrate<float>.
// Measured on B200 (tools/microbench/ffma2_rates.cu, profiles/r2_ffma2_rates.txt): FFMA2 issues every
// 2 cycles per scheduler -- the element rate of scalar FFMA in half the issue slots -- while the packed
// multiply and add (FMUL2 / FADD2) issue only every ~6.6 cycles.  Products and sums are therefore
// written as fused operations that are EXACTLY the plain ones: a*b = fma(a, b, -0) (adding -0 changes
// no value, not even the sign of a zero product) and a+b = fma(a, 1, b).
// ------------------------------------------------------------------------------------------
#ifndef COPTER_F2_PLAIN_MUL_ADD
#define COPTER_F2_PLAIN_MUL_ADD 0     // 1 (A/B knob): FMUL2 / FADD2 for the packed products and sums
#endif
// -0 and 1 as the compiler cannot see them (constant memory can be rewritten by the host, so neither
// nvcc nor ptxas may fold them): with literal constants ptxas turns fma(a, b, -0) straight back into FMUL2
__constant__ float kF2NegZero = -0.0f;
__constant__ float kF2One = 1.0f;
struct F2 {
    float2 v;
    __device__ __forceinline__ F2() {}
    __device__ __forceinline__ explicit F2(float x) { v.x = x; v.y = x; }
    __device__ __forceinline__ F2(float x, float y) { v.x = x; v.y = y; }
};
__device__ __forceinline__ F2 fma_(F2 a, F2 b, F2 c) {
    // inline PTX rather than __ffma2_rn: the compiler must not "simplify" the exact forms below back into mul / add
    F2 r;
    asm("{\n\t.reg .b64 a, b, c, d;\n\t"
        "mov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
        "fma.rn.f32x2 d, a, b, c;\n\t"
        "mov.b64 {%0, %1}, d;\n\t}"
        : "=f"(r.v.x), "=f"(r.v.y) : "f"(a.v.x), "f"(a.v.y), "f"(b.v.x), "f"(b.v.y), "f"(c.v.x), "f"(c.v.y));
    return r;
}
__device__ __forceinline__ F2 operator*(F2 a, F2 b) {
    if (COPTER_F2_PLAIN_MUL_ADD) { F2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
    return fma_(a, b, F2(kF2NegZero));
}
__device__ __forceinline__ F2 operator+(F2 a, F2 b) {
    if (COPTER_F2_PLAIN_MUL_ADD) { F2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
    return fma_(a, F2(kF2One), b);
}
__device__ __forceinline__ F2 operator-(F2 a) { return F2(-a.v.x, -a.v.y); }

// ------------------------------------------------------------------------------------------
// The common case of env_advance as straight-line code, for the K-fused loops: an AIRBORNE env that
// is not touching the ground, is past the first step of its episode (no reset perturbation left)
// and -- fp32 -- has all three angles inside the polynomial range of sincos_poly.  airborne_hot()
// is that precondition; under it airborne_arith() + airborne_flags() give exactly what env_advance gives.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool airborne_hot_c(T z, T dz, T phi, T the, T psi, int st, int steps) {
    // (bitwise on purpose: one straight run of compares, no short-circuit branches)
    int hot = (int)(st == ST_AIRBORNE) & (int)(steps != 1) & (int)!(z > (T)0 && dz > (T)0);
    if constexpr (sizeof(T) == 4) hot &= (int)(fmaxf(fmaxf(fabsf(phi), fabsf(the)), fabsf(psi)) <= 0.78539816f);
    return hot != 0;
}
template <typename T>
__device__ __forceinline__ bool airborne_hot(const T (&s)[12], int st, int steps) {
    return airborne_hot_c<T>(s[4], s[5], s[6], s[8], s[10], st, steps);
}
// A vehicle the airborne arithmetic does not apply to (on the ground, touching it, levelling, crashed): its step is
// the status machine alone -- env_advance without sines, cosines or integration (dynamics_update skips them).
template <typename T>
__device__ __forceinline__ bool on_ground(const T (&s)[12], int st) {
    return st != ST_AIRBORNE || (s[4] > (T)0 && s[5] > (T)0);
}
// env_advance for such a vehicle, written out: dynamics_update (dynamics/__init__.py:147-177) is then the status
// machine alone and _Task.step's tests (task.py:111-130, lander.py:64-72) follow unchanged.  Not for the `direct`
// variants (a LANDED vehicle can take off there).  COPTER_GROUND_FF = 2 steps a grounded lane with this instead of
// the general env_advance.
template <typename T, int VARIANT>
__device__ __forceinline__ void ground_advance(const KParams<T>& kp, T (&s)[12], int& st, int& steps, bool& done, int& cause) {
    using V = Variant<VARIANT>;
    const int st0 = st;                                            // task.py:81 stale status
    if (st0 == ST_LEVELING) { s[6] = (T)0; s[8] = (T)0; st = ST_LANDED; }                                  // :152-156
    else if (st0 == ST_AIRBORNE)                                   // touching the ground: :162-177 (early return, nothing moves)
        st = (s[5] > kp.lvy || abs_t(s[3]) > kp.lvx || abs_t(s[6]) > kp.lang) ? ST_CRASHED : ST_LEVELING;
    cause = 0;
    done = false;
    if (V::lander && st0 == ST_LANDED) {                           // lander.py:64-72
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
    steps = steps < kp.steps_cap ? steps + 1 : kp.steps_cap;       // :130
    if (!done) cause = 0;
}
// the same without the "past the first step" condition: the precondition of airborne_arith_pert()
template <typename T>
__device__ __forceinline__ bool airborne_hot_fresh(const T (&s)[12], int st) {
    return airborne_hot_c<T>(s[4], s[5], s[6], s[8], s[10], st, 0);
}

// Returns the step's ending flags: bit 0 out of bounds, bit 1 over-angle (exclusive: the reference
// tests them with if / elif, task.py:111-118), bit 2 the env's own step limit; 0 = the episode goes
// on.  airborne_cause() turns them into the CAUSE_* bits env_advance reports.
enum { END_OOB = 1, END_ANGLE = 2, END_TIMEOUT = 4 };
// the arithmetic of one such step (no flags, no step counter); L = float, double or F2
template <typename L, typename T>
__device__ __forceinline__ void airborne_arith(const KParams<T>& kp, L (&s)[12], const Forces<L>& f, L& na, L& nc) {
    L sph, cph, sth, cth, sps, cps;
    if constexpr (sizeof(T) == 4) { sincos_poly<L>(s[6], sph, cph); sincos_poly<L>(s[8], sth, cth); sincos_poly<L>(s[10], sps, cps); }
    else sincos3_t(s[6], s[8], s[10], sph, cph, sth, cth, sps, cps);
    const L none[3] = {L((T)0), L((T)0), L((T)0)};
    airborne_integrate<L, T, 3, false>(kp, s, f, none, sph, cph, sth, cth, sps, cps, na, nc);
}
// the same with the reset perturbation `p` of a fresh episode (zeros for an env past its first step): exactly the
// airborne case of dynamics_update, which always carries the perturbation terms
template <typename L, typename T>
__device__ __forceinline__ void airborne_arith_pert(const KParams<T>& kp, L (&s)[12], const Forces<L>& f, const L (&p)[3], L& na, L& nc) {
    L sph, cph, sth, cth, sps, cps;
    if constexpr (sizeof(T) == 4) { sincos_poly<L>(s[6], sph, cph); sincos_poly<L>(s[8], sth, cth); sincos_poly<L>(s[10], sps, cps); }
    else sincos3_t(s[6], s[8], s[10], sph, cph, sth, cth, sps, cps);
    airborne_integrate<L, T, 3, true>(kp, s, f, p, sph, cph, sth, cth, sps, cps, na, nc);
}
// task.py:111-130 with the stale status AIRBORNE, on the state after the step
template <typename T, int VARIANT>
__device__ __forceinline__ int airborne_flags(const KParams<T>& kp, T x, T y, T phi, T the, bool timeout) {
    if constexpr (Variant<VARIANT>::reward == REWARD_TAKEOFF) return timeout ? END_TIMEOUT : 0;
    const bool oob = abs_t(x) >= kp.bounds || abs_t(y) >= kp.bounds;
    const bool ang = !oob && (abs_t(phi) >= kp.max_angle || abs_t(the) >= kp.max_angle);
    return (oob ? END_OOB : 0) | (ang ? END_ANGLE : 0) | (timeout ? END_TIMEOUT : 0);
}
// "Calm" after a straight-line step: the step ended nothing (in bounds, under the angle limit, not
// the step limit) AND the env is still in airborne_hot()'s common case, so the next substep can be
// straight-line too.  Deliberately a little stricter than the two conditions it implies (psi is held
// to the angle limit as well, `<` instead of `<=` at the polynomial range): a lane that is not calm
// just goes back to the exact tests.  Five compares instead of the flags + hot + two warp votes.
template <typename T>
__device__ __forceinline__ bool airborne_calm(const KParams<T>& kp, T x, T y, T z, T dz, T phi, T the, T psi, bool timeout) {
    const T m_xy = max_t(abs_t(x), abs_t(y));
    const T m_ang = max_t(max_t(abs_t(phi), abs_t(the)), abs_t(psi));
    return (int)!timeout & (int)(m_xy < kp.bounds) & (int)(m_ang < kp.calm_angle) & (int)!(z > (T)0 && dz > (T)0);
}
__device__ __forceinline__ int airborne_cause(int end) {
    return ((end & END_OOB) ? CAUSE_OOB : 0) | ((end & END_ANGLE) ? CAUSE_ANGLE : 0) | ((end & END_TIMEOUT) ? CAUSE_TIMEOUT : 0);
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
