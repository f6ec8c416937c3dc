"""
TEST / BASELINE INFRASTRUCTURE ONLY -- single-env, per-object Python port of the reference's
Lander env, in the reference's own execution style: one Python object per env, numpy scalar
arithmetic on a 12-vector, one interpreter-level `step()` call per env-step.

This is what bench.py's `--impl reference` arm and the `cpu_baseline` leg time on the GPU
box's host cores (the reference itself is a Python tree that cannot travel there, and there
is nothing to compile: it has no native code).  It is checked against the executed
reference in tests/test_scalar_port.py, so its throughput is representative of the
reference's own `Lander.step` (~1.3e4 - 1.7e4 steps/s/core).

Restates: gym_copter/dynamics/__init__.py:114-197 (`ScalarCopter.motors`),
gym_copter/envs/task.py:77-137,145-197 and gym_copter/envs/lander.py:46-74 (`ScalarLander`).
"""

import numpy as np

CRASHED, LANDED, LEVELING, AIRBORNE = range(4)


class ScalarCopter:

    def __init__(self, fps=100.0):
        self.B, self.D, self.M, self.L = 5.E-03, 2.E-06, 1.380, 0.350
        self.Ix, self.Iy, self.Iz, self.maxrpm = 2, 2, 3, 15000
        self.G = 9.80665
        self.dt = 1. / fps
        self.s = np.zeros(12)
        self.status = LANDED
        self.kick = np.zeros(6)
        self.ticks = 0

    def place(self, s):
        self.s = np.array(s, dtype=np.float64)
        self.status = AIRBORNE if self.s[4] < 0 else LANDED

    def motors(self, m):
        w2 = (np.array(m) * self.maxrpm * np.pi / 30) ** 2
        u1 = self.B * np.sum(w2)
        u2 = self.L * self.B * ((w2[1] + w2[2]) - (w2[0] + w2[3]))
        u3 = self.L * self.B * ((w2[1] + w2[3]) - (w2[0] + w2[2]))
        u4 = self.D * ((w2[0] + w2[1]) - (w2[2] + w2[3]))
        s = self.s
        cph, cth, cps = np.cos(s[6]), np.cos(s[8]), np.cos(s[10])
        sph, sth, sps = np.sin(s[6]), np.sin(s[8]), np.sin(s[10])
        acc = (-u1 / self.M) * np.array([sph * sps + cph * cps * sth, cph * sps * sth - cps * sph, cph * cth])
        netz = acc[2] + self.G
        if self.status == LANDED and netz < 0:
            self.status = AIRBORNE
        if self.status == LEVELING:
            s[6] = 0
            s[8] = 0
            self.status = LANDED
        elif self.status == AIRBORNE:
            if s[4] > 0 and s[5] > 0:
                hard = s[5] > 1.0 or abs(s[3]) > 2.0 or abs(s[6]) > np.pi / 4
                self.status = CRASHED if hard else LEVELING
                return
            k = self.kick
            d = np.empty(12)
            d[0::2] = s[1::2]
            d[1] = acc[0] + k[0]
            d[3] = acc[1] + k[1]
            d[5] = netz + k[2]
            d[7] = s[11] * s[9] * (self.Iy - self.Iz) / self.Ix + u2 / self.Ix + k[3]
            d[9] = -(s[11] * s[7] * (self.Iz - self.Ix) / self.Iy + u3 / self.Iy) + k[4]
            d[11] = s[9] * s[7] * (self.Ix - self.Iy) / self.Iz + u4 / self.Iz + k[5]
            d[1::2] += k
            self.s = s + self.dt * d
        self.kick = np.zeros(6)
        self.ticks += 1


class ScalarLander:

    def __init__(self, altitude=10.0, max_steps=1000, bounds=10.0, max_angle=np.pi / 4, force=30.0):
        self.altitude, self.max_steps, self.bounds, self.max_angle, self.force = altitude, max_steps, bounds, max_angle, force
        self.rng = np.random.default_rng()

    def _shaping(self, s):
        sh = -(25 * np.sqrt(np.sum(s[0:6] ** 2)) + 50 * np.sqrt(np.sum(s[10:12] ** 2)))
        return sh - 100 if abs(s[5]) > 10 else sh

    def reset(self, force=None):
        self.c = ScalarCopter()
        s0 = np.zeros(12)
        s0[4] = -self.altitude
        self.c.place(s0)
        f = self.rng.uniform(-self.force, self.force, 3) if force is None else np.asarray(force, float)
        self.c.kick = np.concatenate([f, np.zeros(3)]) / self.c.M
        self.prev = self._shaping(self.c.s)
        self.steps = 1
        return self.c.s[:10].astype(np.float32)

    def step(self, action):
        c = self.c
        st0 = c.status
        if st0 != LANDED:
            c.motors(np.clip(action, 0, 1))
        s = c.s
        sh = self._shaping(s)
        reward = sh - self.prev
        self.prev = sh
        done = False
        if st0 == LANDED:
            done = True
            if np.sqrt(s[0] ** 2 + s[2] ** 2) < 2:
                reward += 100
        if abs(s[0]) >= self.bounds or abs(s[2]) >= self.bounds:
            done = True
            reward -= 100
        elif abs(s[6]) >= self.max_angle or abs(s[8]) >= self.max_angle:
            done = True
            reward = -100
        elif st0 == CRASHED:
            done = True
        if self.steps == self.max_steps:
            done = True
        self.steps += 1
        return s[:10].astype(np.float32), reward, done, False, {}


def run_stream(kind, seconds, seed=0):
    """Steps one ScalarLander for ~`seconds` of wall time on the named action stream,
    resetting on done (as a caller of the reference must). Returns (env_steps, elapsed_s)."""
    import time
    rng = np.random.default_rng(seed)
    env = ScalarLander()
    env.rng = rng
    env.reset()
    n, t0 = 0, time.perf_counter()
    while True:
        for _ in range(256):
            if kind == 'const':
                a = 1.625e-2 * np.ones(4)                       # lander.py:21,42
            elif kind == 'randn':
                a = 1.625e-2 * rng.standard_normal(4)           # lander.py:42 --random
            else:
                a = rng.uniform(-1, 1, 4)
            _, _, done, _, _ = env.step(a)
            n += 1
            if done:
                env.reset()
        el = time.perf_counter() - t0
        if el >= seconds:
            return n, el


def _worker(args):
    return run_stream(*args)


def make_pool(procs):
    """A pool of `procs` worker processes that can be reused across run_parallel() calls."""
    import multiprocessing as mp
    return mp.get_context('fork').Pool(procs)


def run_parallel(kind, seconds, procs, pool=None):
    """`procs` independent env loops, one process each. Returns aggregate env-steps/s."""
    own = pool is None
    if own:
        pool = make_pool(procs)
    try:
        res = pool.map(_worker, [(kind, seconds, 1000 + i) for i in range(procs)], chunksize=1)
    finally:
        if own:
            pool.close()
            pool.join()
    return sum(n / el for n, el in res)
