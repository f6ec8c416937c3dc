"""
TEST / BASELINE INFRASTRUCTURE ONLY -- ctypes front end of oracle/copter_oracle.c with the
same reset()/step() surface as oracle/copter_oracle.py's EnvBatch (fp64 only).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .copter_oracle import OracleParams, VARIANTS

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libcopter_oracle.so')
_VARIANT_ID = {v: i for i, v in enumerate(('Lander3D', 'Lander2D', 'Lander1D', 'Hover3D', 'Hover2D', 'Hover1D'))}

_FIELDS = ('B', 'D', 'M', 'L', 'Ix', 'Iy', 'Iz', 'Jr', 'maxrpm', 'landing_vel_x', 'landing_vel_y',
           'landing_angle', 'G', 'fps', 'initial_random_force', 'out_of_bounds_penalty', 'max_angle_deg',
           'bounds', 'initial_altitude', 'target_radius', 'yaw_penalty_factor', 'xyz_penalty_factor',
           'dz_max', 'dz_penalty', 'inside_radius_bonus', 'rho', 'lift_coefficient')


class _CParams(C.Structure):
    _fields_ = [(f, C.c_double) for f in _FIELDS] + [('max_steps', C.c_int32), ('dynamics_model', C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, 'copter_oracle.c')
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(['make', '-s', '-C', HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.oracle_step.restype = C.c_int64
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class CEnvBatch:

    def __init__(self, variant, n, params=None, seed=0, env_offset=0, auto_reset=True, env_ids=None, nthreads=1):
        self.lib = load()
        self.variant, self.vid, self.n = variant, _VARIANT_ID[variant], n
        self.obs_size, self.act_size = len(VARIANTS[variant][1]), VARIANTS[variant][2]
        op = params or OracleParams()
        self.p = _CParams(**{f: float(getattr(op, f)) for f in _FIELDS}, max_steps=int(op.max_steps),
                          dynamics_model=int(op.dynamics_model))
        self.seed, self.auto_reset, self.nthreads = seed, auto_reset, nthreads
        self.env_ids = (np.arange(n, dtype=np.uint64) + np.uint64(env_offset) if env_ids is None
                        else np.ascontiguousarray(env_ids, dtype=np.uint64))
        self.x = np.zeros((n, 12))
        self.status, self.steps, self.episode = (np.zeros(n, np.int32) for _ in range(3))
        self.perturb, self.ticks = np.zeros((n, 6)), np.zeros(n, np.int64)
        self.obs = np.zeros((n, self.obs_size), np.float32)
        self.reward, self.done = np.zeros(n), np.zeros(n, np.uint8)
        self.cause, self.final_steps = np.zeros(n, np.int32), np.zeros(n, np.int32)

    def reset(self, force=None):
        f = None if force is None else np.ascontiguousarray(force, np.float64)
        self.lib.oracle_reset(C.byref(self.p), self.vid, C.c_int64(self.n), _p(self.x), _p(self.status), _p(self.steps),
                              _p(self.episode), _p(self.perturb), _p(self.ticks), _p(self.env_ids),
                              C.c_uint64(self.seed), _p(f), _p(self.obs))
        return self.obs

    def step(self, action, k_substeps=1, force=None):
        a = np.ascontiguousarray(action, np.float64).reshape(self.n, self.act_size)
        f = None if force is None else np.ascontiguousarray(force, np.float64)
        executed = self.lib.oracle_step(
            C.byref(self.p), self.vid, C.c_int64(self.n), _p(self.x), _p(self.status), _p(self.steps), _p(self.episode),
            _p(self.perturb), _p(self.ticks), _p(a), _p(self.env_ids), C.c_uint64(self.seed), _p(f), int(k_substeps),
            int(self.auto_reset), _p(self.obs), _p(self.reward), _p(self.done), _p(self.cause), _p(self.final_steps),
            int(self.nthreads))
        return self.obs, self.reward, self.done.view(np.bool_), {
            'cause': self.cause, 'final_steps': self.final_steps, 'executed': executed}


def throughput(stream, seconds, nthreads, n=1 << 16, seed=0):
    """env-steps/s of the compiled port on `nthreads` host threads (bounded sample)."""
    import time
    env = CEnvBatch('Lander3D', n, seed=seed, nthreads=nthreads)
    env.reset()
    rng = np.random.default_rng(seed)
    pool = []
    for _ in range(4):
        if stream == 'const':
            pool.append(np.full((n, 4), 1.625e-2))
        elif stream == 'randn':
            pool.append(1.625e-2 * rng.standard_normal((n, 4)))
        else:
            pool.append(rng.uniform(-1, 1, (n, 4)))
    env.step(pool[0])
    t0, steps, i = time.perf_counter(), 0, 0
    while time.perf_counter() - t0 < seconds:
        _, _, _, info = env.step(pool[i % 4])
        steps += info['executed']
        i += 1
    return steps / (time.perf_counter() - t0)
