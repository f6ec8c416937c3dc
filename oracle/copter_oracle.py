"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the batched copter step.

A numpy restatement of the reference's algorithm on the hot path, vectorised over N
independent envs.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this; the product package (gym_copter_b200/) never does and has no CPU fallback.

Pinning: the reference ships NO tests or golden vectors of its own, so
this oracle is pinned by EXECUTING the unmodified reference (oracle/refshim.py) in the build
container: tests/test_oracle_vs_reference.py compares them live when /root/reference is
present, and tests/golden/*.npz (made by tests/golden/make_golden.py from the reference)
pin it where the reference tree cannot travel (the GPU box).

Reference lines restated (paths relative to /root/reference):
  gym_copter/dynamics/vehicles/dji_phantom.py:9-26    -> OracleParams vehicle constants
  gym_copter/dynamics/__init__.py:65-76               -> status codes, landing criteria, G
  gym_copter/dynamics/__init__.py:114-197             -> set_motors()
  gym_copter/dynamics/__init__.py:210-217,227-229     -> set_state(), perturb()
  gym_copter/dynamics/__init__.py:249-302,339-350     -> derivative / body-Z rotation
  gym_copter/envs/task.py:77-137                      -> env_step()
  gym_copter/envs/task.py:145-202                     -> reset_where()
  gym_copter/envs/lander.py:17-23,39-74,95-97         -> lander reward / obs / motor map
  attic/gym_copter/envs/{lander2d,lander1d,hover,hover1d,hover2d,hover3d}.py -> VARIANTS

Semantics the batched product adds (the reference is single-env and silent on these; they
are defined in DESIGN.md and mirrored here):
  * same-step auto-reset: when done[i], the flag and the terminal reward are reported and
    env i is replaced by a fresh reset state whose observation is the one returned;
  * k_substeps (frame-skip): K consecutive reference steps under one action, rewards summed,
    an env that finishes at substep j idles for the rest of the call;
  * reset force ~ U(-F, F)^3 from Philox4x32-10, counter (env_lo, env_hi, episode, 0),
    key (seed_lo, seed_hi); the reference draws from the unseedable global numpy RNG
    (envs/task.py:147,199-202) so only the distribution can match, trajectories are compared
    with identical injected forces.
"""

from dataclasses import dataclass

import numpy as np

STATUS_CRASHED, STATUS_LANDED, STATUS_LEVELING, STATUS_AIRBORNE = 0, 1, 2, 3

# name: (kind, obs indices into the 12-state, action size, motor fan-out indices)
VARIANTS = {
    'Lander3D': ('lander', tuple(range(10)), 4, (0, 1, 2, 3)),
    'Lander2D': ('lander', (2, 3, 4, 5, 6, 7), 2, (0, 1, 1, 0)),
    'Lander1D': ('lander', (4, 5), 1, (0, 0, 0, 0)),
    'Hover3D': ('hover', tuple(range(12)), 4, (0, 1, 2, 3)),
    'Hover2D': ('hover', (2, 3, 4, 5, 6, 7), 2, (0, 1, 1, 0)),
    'Hover1D': ('hover', (4, 5), 1, (0, 0, 0, 0)),
}

# attic/gym_copter/envs/takeoff.py:18-91: the 10-component observation, four motor commands handed to
# setMotors UNCLIPPED whatever the flight status, reward = change of -|altitude - target|, never done.
# Kept out of VARIANTS because the golden trajectory files cover the six _Task-shaped variants.
EXTRA_VARIANTS = {'Takeoff': ('takeoff', tuple(range(10)), 4, (0, 1, 2, 3))}
ALL_VARIANTS = dict(VARIANTS, **EXTRA_VARIANTS)

# done-cause bits reported by env_step (non-exclusive; the product's episode statistics)
CAUSE_LANDED, CAUSE_BONUS, CAUSE_OOB, CAUSE_ANGLE, CAUSE_CRASHED, CAUSE_TIMEOUT = 1, 2, 4, 8, 16, 32


@dataclass
class OracleParams:
    # dynamics/vehicles/dji_phantom.py:9-26
    B: float = 5.E-03
    D: float = 2.E-06
    M: float = 1.380
    L: float = 0.350
    Ix: float = 2
    Iy: float = 2
    Iz: float = 3
    Jr: float = 38E-04
    maxrpm: float = 15000
    # dynamics/__init__.py:71-76
    landing_vel_x: float = 2.0
    landing_vel_y: float = 1.0
    landing_angle: float = np.pi / 4
    G: float = 9.80665
    # envs/task.py:25,32-38
    fps: float = 100
    initial_random_force: float = 30
    out_of_bounds_penalty: float = 100
    max_steps: int = 1000
    max_angle_deg: float = 45
    bounds: float = 10
    initial_altitude: float = 10
    # envs/lander.py:17-23
    target_radius: float = 2
    yaw_penalty_factor: float = 50
    xyz_penalty_factor: float = 25
    dz_max: float = 10
    dz_penalty: float = 100
    inside_radius_bonus: float = 100
    # attic/mars/dynamics/__init__.py:83-84,101 and ingenuity.py:55: the alternate vehicle/world model
    rho: float = 1.225
    lift_coefficient: float = 0.4
    dynamics_model: int = 0          # bit 0: lift-model thrust, bit 1: live gyroscopic Omega
    takeoff_target_altitude: float = 5   # attic/gym_copter/envs/takeoff.py:20


# ---------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09 `philox4x32_R(10, ...)`), restated.
# Pinned by the Random123 known-answer vectors in tests/test_philox.py.
# ---------------------------------------------------------------------------------------

_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
_U32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments: uint32 arrays (or scalars) of a common shape. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _U32 for c in (c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _U32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _U32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def reset_force(seed, env_id, episode, scale, dtype=np.float64):
    """
    U(-scale, scale)^3 for (env_id, episode): u32 -> fp64 exactly (u * 2*scale * 2^-32 - scale
    has <= 53 significant bits for scale=30), then ONE rounding to dtype.  Shape [N,3].
    """
    env_id = np.asarray(env_id, dtype=np.uint64)
    r = philox4x32_10(env_id & _U32, env_id >> np.uint64(32),
                      np.asarray(episode, dtype=np.uint64), np.zeros_like(env_id),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    f = [r[j].astype(np.float64) * (2.0 * scale * 2.0 ** -32) - float(scale) for j in range(3)]
    return np.stack(f, axis=-1).astype(dtype)


def source_actions(seed, env_id, step, kind, scale, offset, act_size, dtype=np.float64, tag=1):
    """
    The product's on-device action sources (copter_rollout_*), restated: action_j = offset +
    scale * xi_j with xi from Philox4x32-10, counter (env_lo, env_hi, step, tag), key = seed
    (low word XOR step's high word).  'const' and 'uniform' are bit-exact in `dtype`;
    'randn' (Box-Muller on u1 = (c+1) 2^-32, u2 = c 2^-32, pairs (c0,c1), (c2,c3)) agrees with
    the device to the accuracy of its log / sin / cos.  lander.py:21,42 are the reference's
    two streams: const 1.625e-2 and 1.625e-2 * randn(4).  tag 1 = the action sources, tag 2 = the
    exploration noise of copter_policy_rollout_f32 (kind 'randn').
    """
    T = np.dtype(dtype).type
    env_id = np.asarray(env_id, dtype=np.uint64)
    n = env_id.shape[0]
    xi = np.ones((n, 4), dtype)
    if kind != 'const':
        c = philox4x32_10(env_id & _U32, env_id >> np.uint64(32), np.full(n, step & 0xFFFFFFFF, np.uint64),
                          np.full(n, tag, np.uint64), (seed & 0xFFFFFFFF) ^ (step >> 32), (seed >> 32) & 0xFFFFFFFF)
        c = [x.astype(np.float64) for x in c]
        if kind == 'uniform':
            xi = np.stack([x * 2.0 ** -31 - 1.0 for x in c], -1).astype(dtype)
        else:
            for h in range(2):
                u1, u2 = ((c[2 * h] + 1.0) * 2.0 ** -32).astype(dtype), (c[2 * h + 1] * 2.0 ** -32).astype(dtype)
                r = np.sqrt(T(-2) * np.log(u1))
                ang = T(6.283185307179586) * u2
                xi[:, 2 * h], xi[:, 2 * h + 1] = r * np.cos(ang), r * np.sin(ang)
    return (T(offset) + T(scale) * xi[:, :act_size]).astype(dtype)


# ---------------------------------------------------------------------------------------
# Dynamics (batched restatement of the reference's `Dynamics` object)
# ---------------------------------------------------------------------------------------

class DynamicsBatch:
    """N independent reference `Dynamics` objects (dynamics/__init__.py:33-229) as arrays."""

    def __init__(self, n, params=None, dtype=np.float64):
        self.p = params or OracleParams()
        self.n, self.dtype = n, np.dtype(dtype)
        self.dt = self.dtype.type(1.) / self.dtype.type(self.p.fps)   # :97
        self.x = np.zeros((n, 12), dtype)                              # :101
        self.status = np.full(n, STATUS_LANDED, np.int32)              # :105
        self.ticks = np.zeros(n, np.int64)                             # :98
        self.perturb = np.zeros((n, 6), dtype)                         # :112

    def set_state(self, state, where=None):
        """dynamics/__init__.py:210-217."""
        w = np.ones(self.n, bool) if where is None else where
        self.x[w] = np.asarray(state, self.dtype)[w] if np.ndim(state) == 2 else np.asarray(state, self.dtype)
        self.status[w] = np.where(self.x[w, 4] < 0, STATUS_AIRBORNE, STATUS_LANDED)

    def set_perturb(self, force, where=None):
        """dynamics/__init__.py:227-229 (force: [N,6])."""
        w = np.ones(self.n, bool) if where is None else where
        self.perturb[w] = (np.asarray(force, self.dtype) / self.dtype.type(self.p.M))[w]

    def get_time(self):
        return self.ticks * self.dt                                    # :219-221

    def set_motors(self, motors, where=None):
        """
        dynamics/__init__.py:114-197 for the envs selected by `where` ([N] bool).
        motors: [N,4] in `dtype`.  Operation order follows the reference expression by
        expression so that the fp64 result is bit-comparable.
        """
        p, T = self.p, self.dtype.type
        call = np.ones(self.n, bool) if where is None else where.copy()
        x = self.x
        m = np.asarray(motors, self.dtype)

        # :120-132  motor values -> rad/s -> thrust and torques (Eq. 6)
        om = m * T(p.maxrpm) * T(np.pi) / T(30)
        o = om ** 2
        if p.dynamics_model & 1:
            # attic/mars/dynamics/__init__.py:146-158: rotor lift 0.5 rho S C_L v^2, v = w L/2,
            # S = .05 L 4 (:100); U2/U3 are lift differences without the arm length
            S = T(.05) * T(p.L) * T(4)
            lift = T(0.5) * T(p.rho) * S * T(p.lift_coefficient) * ((om * T(p.L) / T(2)) ** 2)
            U1 = ((lift[:, 0] + lift[:, 1]) + lift[:, 2]) + lift[:, 3]
            U2 = (lift[:, 1] + lift[:, 2]) - (lift[:, 0] + lift[:, 3])
            U3 = (lift[:, 1] + lift[:, 3]) - (lift[:, 0] + lift[:, 2])
        else:
            U1 = T(p.B) * (((o[:, 0] + o[:, 1]) + o[:, 2]) + o[:, 3])
            U2 = T(p.L) * T(p.B) * ((o[:, 1] + o[:, 2]) - (o[:, 0] + o[:, 3]))     # :231-235
            U3 = T(p.L) * T(p.B) * ((o[:, 1] + o[:, 3]) - (o[:, 0] + o[:, 2]))     # :237-241
        U4 = T(p.D) * ((o[:, 0] + o[:, 1]) - (o[:, 2] + o[:, 3]))             # :243-247

        # :139-143  body-Z thrust rotated to NED with the CURRENT angles (:292-302, :339-350)
        phi, the, psi = x[:, 6], x[:, 8], x[:, 10]
        cph, cth, cps = np.cos(phi), np.cos(the), np.cos(psi)
        sph, sth, sps = np.sin(phi), np.sin(the), np.sin(psi)
        bz = -U1 / T(p.M)
        ax = bz * (sph * sps + cph * cps * sth)
        ay = bz * (cph * sps * sth - cps * sph)
        az = bz * (cph * cth)
        netz = az + T(p.G)

        st = self.status
        # :147-149  LANDED -> AIRBORNE when net vertical acceleration is upward
        takeoff = call & (st == STATUS_LANDED) & (netz < 0)
        st[takeoff] = STATUS_AIRBORNE

        # :152-156  LEVELING: zero roll/pitch, become LANDED (falls through to :194-197)
        lev = call & (st == STATUS_LEVELING)
        # :159      AIRBORNE (evaluated as `elif`, so envs that just levelled are excluded)
        air = call & (st == STATUS_AIRBORNE) & ~lev
        x[lev, 6] = 0
        x[lev, 8] = 0
        st[lev] = STATUS_LANDED

        # :162-177  ground contact on the PRE-step state; early return (no integrate, perturb
        #           kept, ticks not incremented).  Note the reference's axis names: "velx" is
        #           dy and "vely" is dz; only phi is angle-tested.
        touch = air & (x[:, 4] > 0) & (x[:, 5] > 0)
        crash = touch & ((x[:, 5] > p.landing_vel_y) | (np.abs(x[:, 3]) > p.landing_vel_x)
                         | (np.abs(x[:, 6]) > p.landing_angle))
        st[crash] = STATUS_CRASHED
        st[touch & ~crash] = STATUS_LEVELING
        integ = air & ~touch

        # :249-290  Eq. 12 (Omega == 0, :135), first perturbation add
        pt = self.perturb
        dphi, dthe, dpsi = x[:, 7], x[:, 9], x[:, 11]
        Ix, Iy, Iz, Jr = T(p.Ix), T(p.Iy), T(p.Iz), T(p.Jr)
        # :135 Omega = 0 in the live model; attic/mars/dynamics/__init__.py:143: u4 of the unsquared speeds
        Omega = ((om[:, 0] + om[:, 1]) - (om[:, 2] + om[:, 3])) if (p.dynamics_model & 2) else T(0)
        d = np.empty_like(x)
        d[:, 0] = x[:, 1]
        d[:, 1] = ax + pt[:, 0]
        d[:, 2] = x[:, 3]
        d[:, 3] = ay + pt[:, 1]
        d[:, 4] = x[:, 5]
        d[:, 5] = netz + pt[:, 2]
        d[:, 6] = dphi
        d[:, 7] = dpsi * dthe * (Iy - Iz) / Ix - Jr / Ix * dthe * Omega + U2 / Ix + pt[:, 3]
        d[:, 8] = dthe
        d[:, 9] = -(dpsi * dphi * (Iz - Ix) / Iy + Jr / Iy * dphi * Omega + U3 / Iy) + pt[:, 4]
        d[:, 10] = dpsi
        d[:, 11] = dthe * dphi * (Ix - Iy) / Iz + U4 / Iz + pt[:, 5]
        # :183  the perturbation is added a SECOND time
        d[:, 1::2] += pt
        # :187  forward Euler, all derivatives from the old state
        x[integ] = (x + self.dt * d)[integ]

        # :194-197  clear perturbation, advance time (skipped by the :177 early return)
        fin = call & ~touch
        pt[fin] = 0
        self.ticks[fin] += 1


# ---------------------------------------------------------------------------------------
# Env (batched restatement of `_Task` + `Lander` / hover hooks, plus auto-reset + K-fusion)
# ---------------------------------------------------------------------------------------

class EnvBatch:

    def __init__(self, variant, n, params=None, dtype=np.float64, seed=0, env_offset=0,
                 auto_reset=True, env_ids=None):
        self.kind, self.obs_idx, self.act_size, self.fanout = ALL_VARIANTS[variant]
        self.obs_idx, self.fanout = list(self.obs_idx), list(self.fanout)
        self.variant, self.n = variant, n
        self.p = params or OracleParams()
        self.dtype = np.dtype(dtype)
        self.seed, self.env_offset, self.auto_reset = seed, env_offset, auto_reset
        # global env ids (key the Philox stream); default: a contiguous shard
        self.env_ids = (np.arange(n, dtype=np.uint64) + np.uint64(env_offset) if env_ids is None
                        else np.asarray(env_ids, dtype=np.uint64))
        self.dyn = DynamicsBatch(n, self.p, dtype)
        self.steps = np.zeros(n, np.int64)
        self.episode = np.zeros(n, np.int64)
        self.max_angle = np.radians(self.p.max_angle_deg)          # envs/task.py:58
        self.ep_mask = 0x7FFFF                                     # the product's compact episode field (wide: 2^32 - 1)

    # ---- helpers ----------------------------------------------------------------------

    def _shaping(self, x):
        """envs/lander.py:48-56."""
        T, p = self.dtype.type, self.p
        q = x[:, 0:6] ** 2
        spos = ((((q[:, 0] + q[:, 1]) + q[:, 2]) + q[:, 3]) + q[:, 4]) + q[:, 5]
        spsi = x[:, 10] ** 2 + x[:, 11] ** 2
        sh = -(T(p.xyz_penalty_factor) * np.sqrt(spos) + T(p.yaw_penalty_factor) * np.sqrt(spsi))
        return np.where(np.abs(x[:, 5]) > p.dz_max, sh - T(p.dz_penalty), sh)

    def observe(self):
        """envs/task.py:133 + the variant's _get_state: float32 cast of the projected state."""
        return self.dyn.x[:, self.obs_idx].astype(np.float32)

    def forces_for(self, env_mask):
        return reset_force(self.seed, self.env_ids, self.episode, self.p.initial_random_force, self.dtype)

    def reset_where(self, w, force=None):
        """
        envs/task.py:145-197 for the envs in `w`: default pose (0,0,-alt), fresh dynamics,
        perturbation = force/M, steps = 1 after the priming step (which touches nothing else:
        task.py:93 skips setMotors, reward 0, prev_shaping := shaping(s0)).
        `force` [N,3] overrides the Philox draw (parity with injected forces).
        """
        if not np.any(w):
            return
        s0 = np.zeros(12, self.dtype)
        s0[4] = -self.dtype.type(self.p.initial_altitude)
        self.dyn.set_state(s0, w)
        self.dyn.ticks[w] = 0
        f3 = self.forces_for(w) if force is None else np.asarray(force, self.dtype)
        f6 = np.zeros((self.n, 6), self.dtype)
        f6[:, 0:3] = f3
        self.dyn.set_perturb(f6, w)
        self.steps[w] = 1

    def reset(self, force=None, keep_episode=None):
        """keep_episode=None mirrors the product's host shell: the first reset() starts every env at
        episode 0, every later one moves each env on to its next episode index, so that a reset() per
        episode draws a NEW force each time like the reference's np.random.uniform (task.py:175-184)."""
        keep = getattr(self, '_ever_reset', False) if keep_episode is None else keep_episode
        if keep:
            self.episode[:] = (self.episode + 1) & self.ep_mask
        else:
            self.episode[:] = 0
        self._ever_reset = True
        self.reset_where(np.ones(self.n, bool), force)
        return self.observe()

    # ---- one reference step -------------------------------------------------------------

    def _single_step(self, action, live):
        """
        envs/task.py:77-137 for the envs in `live`.  Returns (reward, done, cause) over all N
        (zeros / False outside `live`).
        """
        T, p, d = self.dtype.type, self.p, self.dyn
        st0 = d.status.copy()                                      # :81 (stale status)
        a = np.asarray(action, self.dtype).reshape(self.n, self.act_size)
        cause = np.zeros(self.n, np.int32)
        if self.kind == 'takeoff':                                 # attic takeoff.py:57-88
            pre_t = -np.abs(-d.x[:, 4] - T(p.takeoff_target_altitude))
            d.set_motors(a[:, self.fanout], live)                  # no clip (:64), whatever the status
            reward = -np.abs(-d.x[:, 4] - T(p.takeoff_target_altitude)) - pre_t       # :77-86
            timeout = self.steps == p.max_steps                    # the batched step limit (not in the attic env)
            done = live & timeout
            cause |= timeout * CAUSE_TIMEOUT
            self.steps[live] += 1
            return np.where(live, reward, T(0)), done, np.where(done, cause, 0)
        motors = np.clip(a, 0, 1)[:, self.fanout]                  # :91 + _get_motors
        pre = self._shaping(d.x)                                   # == prev_shaping
        d.set_motors(motors, live & (st0 != STATUS_LANDED))        # :86-94
        x = d.x
        if self.kind == 'lander':
            reward = self._shaping(x) - pre                        # lander.py:58-62
            landed = live & (st0 == STATUS_LANDED)                 # lander.py:64-72
            bonus = landed & (np.sqrt(x[:, 0] ** 2 + x[:, 2] ** 2) < p.target_radius)
            reward = np.where(bonus, reward + T(p.inside_radius_bonus), reward)
            done = landed.copy()
            cause |= landed * CAUSE_LANDED | bonus * CAUSE_BONUS
        else:
            reward = np.ones(self.n, self.dtype)                   # attic hover.py:18-21
            done = np.zeros(self.n, bool)
        oob = (np.abs(x[:, 0]) >= p.bounds) | (np.abs(x[:, 2]) >= p.bounds)          # :111
        ang = ~oob & ((np.abs(x[:, 6]) >= self.max_angle) | (np.abs(x[:, 8]) >= self.max_angle))  # :116
        crashed = ~oob & ~ang & (st0 == STATUS_CRASHED)                              # :121
        reward = np.where(oob, reward - T(p.out_of_bounds_penalty), reward)
        reward = np.where(ang, -T(p.out_of_bounds_penalty), reward)
        timeout = self.steps == p.max_steps                                          # :128
        done = live & (done | oob | ang | crashed | timeout)
        cause |= (oob * CAUSE_OOB | ang * CAUSE_ANGLE | (st0 == STATUS_CRASHED) * CAUSE_CRASHED
                  | timeout * CAUSE_TIMEOUT)
        self.steps[live] += 1                                                        # :130
        return np.where(live, reward, T(0)), done, np.where(done, cause, 0)

    # ---- batched step with K-fusion and same-step auto-reset ----------------------------

    def step(self, action, k_substeps=1, force=None):
        """
        Returns (obs f32 [N,O], reward [N], done bool [N], info) where info carries
        'steps_taken' (substeps actually executed), 'final_steps' (the episode length counter
        at termination) and 'cause'.
        """
        n = self.n
        total = np.zeros(n, self.dtype)
        done_any = np.zeros(n, bool)
        cause_any = np.zeros(n, np.int32)
        taken = np.zeros(n, np.int64)
        final_steps = np.zeros(n, np.int64)
        for _ in range(k_substeps):
            live = ~done_any
            r, dn, cs = self._single_step(action, live)
            total = total + r
            taken += live
            final_steps = np.where(dn, self.steps, final_steps)
            cause_any |= cs
            done_any |= dn
            if self.auto_reset and np.any(dn):
                self.episode[dn] = (self.episode[dn] + 1) & self.ep_mask
                self.reset_where(dn, force)
        return self.observe(), total, done_any, {
            'steps_taken': taken, 'final_steps': final_steps, 'cause': cause_any}
