// TEST INFRASTRUCTURE ONLY -- the CPU restatement of the KERNEL arithmetic.
//
// This file instantiates gym_copter_b200/csrc/copter_core.h -- the one header the CUDA kernels take
// their arithmetic from -- with a plain host compiler (g++ -ffp-contract=off, see oracle/Makefile) and
// loops it over a batch of envs, one env at a time, through env_launch(): k substeps under one action
// through the GENERAL env_advance, no fast paths, no warp votes, no packed pairs.  It serves two checks:
//   * on the CPU (tests/test_host_restatement.py): the fp64 instantiation against the numpy oracle that
//     is pinned to the executed reference (<= 1e-9, discrete outputs exact), and the fp32 instantiation
//     against the same oracle (<= 1e-4, with the rate of threshold flips measured per action stream);
//   * on the GPU (tests/test_gpu_host_exact.py): the fp32 kernels against the fp32 instantiation --
//     state, flight status, step counters, episode indices and done flags BIT FOR BIT, zero flips.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library; the product
// package never does.  The reference lines each function restates are cited in copter_core.h.

#include <stdint.h>
#include <string.h>

#include "../gym_copter_b200/csrc/copter_core.h"

using namespace copter;

namespace {

template <typename T, int VARIANT>
int64_t step_batch(const KParams<T>& kp, int64_t n, T* state, int32_t* status, int32_t* steps, uint32_t* episode,
                   const T* action, const uint64_t* env_ids, uint64_t seed, const T* init_force, int k, int auto_reset,
                   float* obs, T* reward, uint8_t* done, uint8_t* cause, int32_t* executed, float* final_obs, int32_t* final_steps) {
    constexpr int O = Variant<VARIANT>::O, A = Variant<VARIANT>::A, FIRST = Variant<VARIANT>::first;
    int64_t total_executed = 0;
    for (int64_t i = 0; i < n; ++i) {
        T s[12], m[4], pert[3] = {(T)0, (T)0, (T)0};
        for (int j = 0; j < 12; ++j) s[j] = state[12 * i + j];
        int st = status[i], stp = steps[i];
        uint32_t ep = episode[i];
        motors_from_action<T, VARIANT>(action + (int64_t)A * i, m);
        if (stp == 1) {                                     // the reset perturbation of a fresh episode
            T f[3];
            if (init_force) { f[0] = init_force[3 * i]; f[1] = init_force[3 * i + 1]; f[2] = init_force[3 * i + 2]; }
            else reset_force<T>(kp, seed, env_ids[i], ep, f);
            for (int j = 0; j < 3; ++j) pert[j] = f[j] * kp.invM;
        }
        EnvOut<T> out;
        env_launch<T, VARIANT>(kp, s, st, stp, ep, m, pert, k, auto_reset != 0, out);
        for (int j = 0; j < 12; ++j) state[12 * i + j] = s[j];
        status[i] = st; steps[i] = stp; episode[i] = ep;
        reward[i] = out.reward; done[i] = out.done ? 1 : 0;
        if (cause) cause[i] = (uint8_t)out.cause;
        if (executed) executed[i] = out.executed;
        if (final_steps) final_steps[i] = out.final_steps;
        total_executed += out.executed;
        if (obs) for (int j = 0; j < O; ++j) obs[(int64_t)O * i + j] = (float)s[FIRST + j];
        if (final_obs && out.has_final) for (int j = 0; j < O; ++j) final_obs[(int64_t)O * i + j] = (float)out.final_state[FIRST + j];
    }
    return total_executed;
}

template <typename T>
int64_t step_dispatch(const CopterParams* p, int variant, int wide, int64_t n, T* state, int32_t* status, int32_t* steps,
                      uint32_t* episode, const T* action, const uint64_t* env_ids, uint64_t seed, const T* init_force, int k,
                      int auto_reset, float* obs, T* reward, uint8_t* done, uint8_t* cause, int32_t* executed, float* final_obs, int32_t* final_steps) {
    if (p->max_steps < 1 || p->max_steps > (wide ? COPTER_MAX_STEPS_LIMIT_WIDE : COPTER_MAX_STEPS_LIMIT)) return -1;
    const KParams<T> kp = make_kparams<T>(*p, wide != 0);
#define COPTER_HOST_CASE(V) case V: return step_batch<T, V>(kp, n, state, status, steps, episode, action, env_ids, seed, init_force, k, auto_reset, obs, reward, done, cause, executed, final_obs, final_steps)
    switch (variant) {
        COPTER_HOST_CASE(COPTER_LANDER3D); COPTER_HOST_CASE(COPTER_LANDER2D); COPTER_HOST_CASE(COPTER_LANDER1D);
        COPTER_HOST_CASE(COPTER_HOVER3D);  COPTER_HOST_CASE(COPTER_HOVER2D);  COPTER_HOST_CASE(COPTER_HOVER1D);
        COPTER_HOST_CASE(COPTER_TAKEOFF);
        default: return -1;
    }
#undef COPTER_HOST_CASE
}

template <typename T>
void reset_batch(const CopterParams* p, int variant, int wide, int64_t n, T* state, int32_t* status, int32_t* steps,
                 uint32_t* episode, int keep_episode, float* obs) {
    const KParams<T> kp = make_kparams<T>(*p, wide != 0);
    static const int O_[COPTER_NUM_VARIANTS] = {10, 6, 2, 12, 6, 2, 10}, F_[COPTER_NUM_VARIANTS] = {0, 2, 4, 0, 2, 4, 0};
    const int O = O_[variant], FIRST = F_[variant];
    for (int64_t i = 0; i < n; ++i) {
        T s[12]; int st, stp;
        reset_state<T>(kp, s, st, stp);
        for (int j = 0; j < 12; ++j) state[12 * i + j] = s[j];
        status[i] = st; steps[i] = stp;
        episode[i] = keep_episode ? ((episode[i] + 1) & kp.ep_mask) : 0u;
        if (obs) for (int j = 0; j < O; ++j) obs[(int64_t)O * i + j] = (float)s[FIRST + j];
    }
}

// Dynamics.setMotors driven directly (dynamics/__init__.py:114-197): the restatement of copter_dynamics_kernel
template <typename T>
void dynamics_batch(const CopterParams* p, int64_t n, T* state, uint8_t* status, int32_t* ticks, T* perturb, const T* motors) {
    const KParams<T> kp = make_kparams<T>(*p);
    for (int64_t i = 0; i < n; ++i) {
        T s[12], pt[6];
        for (int j = 0; j < 12; ++j) s[j] = state[12 * i + j];
        for (int j = 0; j < 6; ++j) pt[j] = perturb[6 * i + j];
        int st = status[i];
        const Forces<T> f = motor_forces<T>(kp, motors[4 * i], motors[4 * i + 1], motors[4 * i + 2], motors[4 * i + 3]);
        T na, nc;
        const bool finished = dynamics_update<T, 6, true>(kp, s, st, f, pt, na, nc);
        for (int j = 0; j < 12; ++j) state[12 * i + j] = s[j];
        status[i] = (uint8_t)st;
        if (finished) { for (int j = 0; j < 6; ++j) perturb[6 * i + j] = (T)0; ticks[i] += 1; }
    }
}

}  // namespace

extern "C" {

int copter_host_abi(void) { return COPTER_ABI_VERSION; }

int64_t copter_host_step_f32(const CopterParams* p, int variant, int wide, int64_t n, float* state, int32_t* status, int32_t* steps,
                             uint32_t* episode, const float* action, const uint64_t* env_ids, uint64_t seed, const float* init_force,
                             int k, int auto_reset, float* obs, float* reward, uint8_t* done, uint8_t* cause, int32_t* executed,
                             float* final_obs, int32_t* final_steps) {
    return step_dispatch<float>(p, variant, wide, n, state, status, steps, episode, action, env_ids, seed, init_force, k, auto_reset,
                                obs, reward, done, cause, executed, final_obs, final_steps);
}
int64_t copter_host_step_f64(const CopterParams* p, int variant, int wide, int64_t n, double* state, int32_t* status, int32_t* steps,
                             uint32_t* episode, const double* action, const uint64_t* env_ids, uint64_t seed, const double* init_force,
                             int k, int auto_reset, float* obs, double* reward, uint8_t* done, uint8_t* cause, int32_t* executed,
                             float* final_obs, int32_t* final_steps) {
    return step_dispatch<double>(p, variant, wide, n, state, status, steps, episode, action, env_ids, seed, init_force, k, auto_reset,
                                 obs, reward, done, cause, executed, final_obs, final_steps);
}
void copter_host_reset_f32(const CopterParams* p, int variant, int wide, int64_t n, float* state, int32_t* status, int32_t* steps,
                           uint32_t* episode, int keep_episode, float* obs) {
    reset_batch<float>(p, variant, wide, n, state, status, steps, episode, keep_episode, obs);
}
void copter_host_reset_f64(const CopterParams* p, int variant, int wide, int64_t n, double* state, int32_t* status, int32_t* steps,
                           uint32_t* episode, int keep_episode, float* obs) {
    reset_batch<double>(p, variant, wide, n, state, status, steps, episode, keep_episode, obs);
}
void copter_host_dynamics_f32(const CopterParams* p, int64_t n, float* state, uint8_t* status, int32_t* ticks, float* perturb, const float* motors) {
    dynamics_batch<float>(p, n, state, status, ticks, perturb, motors);
}
void copter_host_dynamics_f64(const CopterParams* p, int64_t n, double* state, uint8_t* status, int32_t* ticks, double* perturb, const double* motors) {
    dynamics_batch<double>(p, n, state, status, ticks, perturb, motors);
}
// the fp32 sin / cos of the kernels, for the accuracy test of the large-angle reduction
void copter_host_sincos_f32(const float* a, int64_t n, float* s, float* c) {
    for (int64_t i = 0; i < n; ++i) sincos_t(a[i], s[i], c[i]);
}
void copter_host_reset_force_f32(const CopterParams* p, int64_t n, const uint64_t* env_ids, const uint32_t* episode, uint64_t seed, float* out) {
    const KParams<float> kp = make_kparams<float>(*p);
    for (int64_t i = 0; i < n; ++i) { float f[3]; reset_force<float>(kp, seed, env_ids[i], episode[i], f); memcpy(out + 3 * i, f, sizeof f); }
}

}  // extern "C"
