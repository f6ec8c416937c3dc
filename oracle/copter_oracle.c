/*
 * TEST / BASELINE INFRASTRUCTURE ONLY -- plain-C restatement of the reference's copter step
 * (fp64, the reference's precision), one env per loop iteration, optional OpenMP over envs.
 * Nothing in the product package links or loads this file.
 *
 * Used (a) as a second, independently written checker for large parity cases
 * (tests/test_c_oracle.py pins it against the numpy oracle and the golden vectors recorded
 * from the executed reference) and (b) as the "compiled CPU port" context number beside
 * bench.py's cpu_baseline.
 *
 * Restates, function by function (paths relative to /root/reference):
 *   set_motors()   gym_copter/dynamics/__init__.py:114-197 (+ :231-247, :249-302, :339-350)
 *   shaping()      gym_copter/envs/lander.py:48-56
 *   single_step()  gym_copter/envs/task.py:77-137, gym_copter/envs/lander.py:58-72,
 *                  attic/gym_copter/envs/hover.py:18-21
 *   reset_env()    gym_copter/envs/task.py:145-197, gym_copter/dynamics/__init__.py:210-229
 *   philox()       Philox4x32-10, Random123 v1.09 (Salmon et al. SC'11)
 * and the product-defined batched semantics (same-step auto-reset, K-substep frame skip with
 * idle-after-done, Philox-keyed reset force) exactly as oracle/copter_oracle.py states them.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { CRASHED = 0, LANDED = 1, LEVELING = 2, AIRBORNE = 3 };
enum { C_LANDED = 1, C_BONUS = 2, C_OOB = 4, C_ANGLE = 8, C_CRASHED = 16, C_TIMEOUT = 32 };

typedef struct {
    double B, D, M, L, Ix, Iy, Iz, Jr, maxrpm;
    double landing_vel_x, landing_vel_y, landing_angle, G;
    double fps, initial_random_force, out_of_bounds_penalty, max_angle_deg, bounds, initial_altitude;
    double target_radius, yaw_penalty_factor, xyz_penalty_factor, dz_max, dz_penalty, inside_radius_bonus;
    double rho, lift_coefficient;
    int32_t max_steps, dynamics_model;       /* bit 0: lift-model thrust, bit 1: live gyroscopic Omega (attic/mars) */
} OracleParams;

/* variant tables: Lander3D, Lander2D, Lander1D, Hover3D, Hover2D, Hover1D */
static const int V_OBS[6] = {10, 6, 2, 12, 6, 2};
static const int V_ACT[6] = {4, 2, 1, 4, 2, 1};
static const int V_FIRST[6] = {0, 2, 4, 0, 2, 4};
static const int V_LANDER[6] = {1, 1, 1, 0, 0, 0};
static const int V_FAN[6][4] = {{0, 1, 2, 3}, {0, 1, 1, 0}, {0, 0, 0, 0}, {0, 1, 2, 3}, {0, 1, 1, 0}, {0, 0, 0, 0}};

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_reset_force(uint64_t seed, uint64_t env, uint32_t episode, double scale, double f[3]) {
    const uint32_t ctr[4] = {(uint32_t)env, (uint32_t)(env >> 32), episode, 0u};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    oracle_philox4x32_10(ctr, key, r);
    for (int j = 0; j < 3; ++j) f[j] = (double)r[j] * (2.0 * scale * 0x1p-32) - scale;
}

/* dynamics/__init__.py:114-197.  Returns 1 when the call reached :194-197 (perturbation
   cleared, ticks += 1), 0 on the :177 early return. */
static int set_motors(const OracleParams* p, double* x, int32_t* status, double* pt, const double m[4]) {
    double o[4], w[4], U1, U2, U3;
    for (int j = 0; j < 4; ++j) { w[j] = m[j] * p->maxrpm * M_PI / 30; o[j] = w[j] * w[j]; }   /* :120-124 */
    if (p->dynamics_model & 1) {          /* attic/mars/dynamics/__init__.py:146-158 */
        double lift[4];
        const double S = .05 * p->L * 4;
        for (int j = 0; j < 4; ++j) { const double v = w[j] * p->L / 2; lift[j] = 0.5 * p->rho * S * p->lift_coefficient * (v * v); }
        U1 = ((lift[0] + lift[1]) + lift[2]) + lift[3];
        U2 = (lift[1] + lift[2]) - (lift[0] + lift[3]);
        U3 = (lift[1] + lift[3]) - (lift[0] + lift[2]);
    } else {
        U1 = p->B * (((o[0] + o[1]) + o[2]) + o[3]);
        U2 = p->L * p->B * ((o[1] + o[2]) - (o[0] + o[3]));
        U3 = p->L * p->B * ((o[1] + o[3]) - (o[0] + o[2]));
    }
    const double U4 = p->D * ((o[0] + o[1]) - (o[2] + o[3]));
    const double cph = cos(x[6]), cth = cos(x[8]), cps = cos(x[10]);
    const double sph = sin(x[6]), sth = sin(x[8]), sps = sin(x[10]);
    const double bz = -U1 / p->M;
    const double ax = bz * (sph * sps + cph * cps * sth), ay = bz * (cph * sps * sth - cps * sph), az = bz * (cph * cth);
    const double netz = az + p->G;
    if (*status == LANDED && netz < 0) *status = AIRBORNE;                   /* :147-149 */
    if (*status == LEVELING) {                                               /* :152-156 */
        x[6] = 0; x[8] = 0; *status = LANDED;
    } else if (*status == AIRBORNE) {
        if (x[4] > 0 && x[5] > 0) {                                          /* :162-177 */
            *status = (x[5] > p->landing_vel_y || fabs(x[3]) > p->landing_vel_x || fabs(x[6]) > p->landing_angle) ? CRASHED : LEVELING;
            return 0;
        }
        const double dphi = x[7], dthe = x[9], dpsi = x[11];
        const double Omega = (p->dynamics_model & 2) ? (w[0] + w[1]) - (w[2] + w[3]) : 0;   /* :135 vs attic/mars :143 */
        double d[12];
        d[0] = x[1]; d[1] = ax + pt[0]; d[2] = x[3]; d[3] = ay + pt[1]; d[4] = x[5]; d[5] = netz + pt[2];
        d[6] = dphi;
        d[7] = dpsi * dthe * (p->Iy - p->Iz) / p->Ix - p->Jr / p->Ix * dthe * Omega + U2 / p->Ix + pt[3];
        d[8] = dthe;
        d[9] = -(dpsi * dphi * (p->Iz - p->Ix) / p->Iy + p->Jr / p->Iy * dphi * Omega + U3 / p->Iy) + pt[4];
        d[10] = dpsi;
        d[11] = dthe * dphi * (p->Ix - p->Iy) / p->Iz + U4 / p->Iz + pt[5];
        for (int j = 0; j < 6; ++j) d[2 * j + 1] += pt[j];                    /* :183 */
        const double dt = 1. / p->fps;
        for (int j = 0; j < 12; ++j) x[j] += dt * d[j];                       /* :187 */
    }
    for (int j = 0; j < 6; ++j) pt[j] = 0;                                    /* :194 */
    return 1;
}

static double shaping(const OracleParams* p, const double* x) {              /* lander.py:48-56 */
    double spos = 0;
    for (int j = 0; j < 6; ++j) spos += x[j] * x[j];
    const double spsi = x[10] * x[10] + x[11] * x[11];
    double sh = -(p->xyz_penalty_factor * sqrt(spos) + p->yaw_penalty_factor * sqrt(spsi));
    if (fabs(x[5]) > p->dz_max) sh -= p->dz_penalty;
    return sh;
}

static void reset_env(const OracleParams* p, double* x, int32_t* status, int32_t* steps, double* pt, int64_t* ticks,
                      const double f[3]) {
    memset(x, 0, 12 * sizeof(double));
    x[4] = -p->initial_altitude;                                             /* task.py:164-171 */
    *status = x[4] < 0 ? AIRBORNE : LANDED;                                   /* dynamics:215-217 */
    for (int j = 0; j < 6; ++j) pt[j] = j < 3 ? f[j] / p->M : 0.0;            /* task.py:179-186, dynamics:229 */
    *steps = 1;                                                               /* task.py:191,197 */
    *ticks = 0;
}

static void single_step(const OracleParams* p, int variant, double* x, int32_t* status, int32_t* steps, double* pt,
                        int64_t* ticks, const double* action, double* reward, int* done, int* cause) {
    const int st0 = *status;                                                  /* task.py:81 */
    const double pre = shaping(p, x);
    if (st0 != LANDED) {                                                      /* :86-94 */
        double m[4];
        for (int j = 0; j < 4; ++j) { const double a = action[V_FAN[variant][j]]; m[j] = a < 0 ? 0 : (a > 1 ? 1 : a); }
        if (set_motors(p, x, status, pt, m)) *ticks += 1;
    }
    double r; int dn = 0, cs = 0;
    if (V_LANDER[variant]) {
        r = shaping(p, x) - pre;                                              /* lander.py:58-62 */
        if (st0 == LANDED) {                                                  /* :64-72 */
            dn = 1; cs |= C_LANDED;
            if (sqrt(x[0] * x[0] + x[2] * x[2]) < p->target_radius) { r += p->inside_radius_bonus; cs |= C_BONUS; }
        }
    } else {
        r = 1;
    }
    const double max_angle = p->max_angle_deg * M_PI / 180.0;                 /* np.radians, task.py:58 */
    if (fabs(x[0]) >= p->bounds || fabs(x[2]) >= p->bounds) { dn = 1; r -= p->out_of_bounds_penalty; cs |= C_OOB; }
    else if (fabs(x[6]) >= max_angle || fabs(x[8]) >= max_angle) { dn = 1; r = -p->out_of_bounds_penalty; cs |= C_ANGLE; }
    else if (st0 == CRASHED) dn = 1;
    if (st0 == CRASHED) cs |= C_CRASHED;
    if (*steps == p->max_steps) { dn = 1; cs |= C_TIMEOUT; }                  /* :128 */
    *steps += 1;
    *reward = r; *done = dn; *cause = dn ? cs : 0;
}

/* Reset all n envs (episode := 0). force: [n][3] or NULL for the Philox draw. */
void oracle_reset(const OracleParams* p, int variant, int64_t n, double* x, int32_t* status, int32_t* steps,
                  int32_t* episode, double* perturb, int64_t* ticks, const uint64_t* env_ids, uint64_t seed,
                  const double* force, float* obs) {
    for (int64_t i = 0; i < n; ++i) {
        double f[3];
        episode[i] = 0;
        if (force) memcpy(f, force + 3 * i, sizeof f); else oracle_reset_force(seed, env_ids[i], 0, p->initial_random_force, f);
        reset_env(p, x + 12 * i, status + i, steps + i, perturb + 6 * i, ticks + i, f);
        if (obs) for (int j = 0; j < V_OBS[variant]; ++j) obs[i * V_OBS[variant] + j] = (float)x[12 * i + V_FIRST[variant] + j];
    }
}

/* k reference steps per env under one action; rewards summed; idle after done; same-step
   auto-reset when auto_reset != 0.  Returns the number of env-steps executed. */
int64_t oracle_step(const OracleParams* p, int variant, int64_t n, double* x, int32_t* status, int32_t* steps,
                    int32_t* episode, double* perturb, int64_t* ticks, const double* action, const uint64_t* env_ids,
                    uint64_t seed, const double* force, int k, int auto_reset, float* obs, double* reward,
                    uint8_t* done, int32_t* cause, int32_t* final_steps, int nthreads) {
    const int A = V_ACT[variant], O = V_OBS[variant], first = V_FIRST[variant];
    int64_t executed = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : executed)
#endif
    for (int64_t i = 0; i < n; ++i) {
        double total = 0; int dn_any = 0, cs_any = 0, fs = 0;
        for (int s = 0; s < k && !dn_any; ++s) {
            double r; int dn, cs;
            single_step(p, variant, x + 12 * i, status + i, steps + i, perturb + 6 * i, ticks + i, action + A * i, &r, &dn, &cs);
            total += r; ++executed;
            if (dn) {
                dn_any = 1; cs_any = cs; fs = steps[i];
                if (auto_reset) {
                    double f[3];
                    episode[i] += 1;
                    if (force) memcpy(f, force + 3 * i, sizeof f);
                    else oracle_reset_force(seed, env_ids[i], (uint32_t)episode[i], p->initial_random_force, f);
                    reset_env(p, x + 12 * i, status + i, steps + i, perturb + 6 * i, ticks + i, f);
                }
            }
        }
        reward[i] = total; done[i] = (uint8_t)dn_any;
        if (cause) cause[i] = cs_any;
        if (final_steps) final_steps[i] = fs;
        if (obs) for (int j = 0; j < O; ++j) obs[i * O + j] = (float)x[12 * i + first + j];
    }
    return executed;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
