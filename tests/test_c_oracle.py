"""The plain-C restatement (oracle/copter_oracle.c) against the numpy oracle and the golden
vectors recorded from the executed reference.  glibc and numpy sin/cos may differ in the last
bit, so floats are compared at 1e-12; flags, step counters and Philox draws bit-exact."""
import os

import numpy as np
import pytest

from oracle.c_oracle import CEnvBatch, load
from oracle.copter_oracle import EnvBatch, VARIANTS, reset_force


def merr(a, ref):
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), 1.0)))


def test_philox_and_force_bit_exact():
    import ctypes as C
    lib = load()
    f = (C.c_double * 3)()
    for env, ep, seed in ((0, 0, 0), (12345678901234, 77, 0xABCDEF0123456789), (2**32 - 1, 2**19 - 1, 5)):
        lib.oracle_reset_force(C.c_uint64(seed), C.c_uint64(env), C.c_uint32(ep), C.c_double(30.0), f)
        assert list(f) == list(reset_force(seed, [env], [ep], 30.0)[0])


@pytest.mark.parametrize('variant', list(VARIANTS))
def test_golden_trajectories(golden_dir, variant):
    g = np.load(os.path.join(golden_dir, 'traj_%s.npz' % variant))
    act = g['action']
    T, N, A = act.shape
    env = CEnvBatch(variant, N, seed=int(g['seed']))
    env.reset()
    for t in range(T):
        obs, r, done, info = env.step(act[t].astype(np.float64))
        assert np.array_equal(done, g['done'][t]) and merr(r, g['reward'][t]) <= 1e-12
        if t % 10 == 9:
            assert merr(env.x, g['state_every10'][t // 10]) <= 1e-12


@pytest.mark.parametrize('k,nthreads', [(1, 1), (4, 2)])
def test_vs_numpy_oracle(k, nthreads):
    N, T = 3000, 250
    rng = np.random.default_rng(k)
    a_env, b_env = EnvBatch('Lander3D', N, seed=9, env_offset=10**10), CEnvBatch('Lander3D', N, seed=9, env_offset=10**10, nthreads=nthreads)
    assert np.array_equal(a_env.reset(), b_env.reset())
    n_done = 0
    for t in range(T):
        a = np.where(np.arange(N)[:, None] % 3 == 0, rng.uniform(-1, 1, (N, 4)), 1.625e-2 * rng.standard_normal((N, 4)))
        o1, r1, d1, i1 = a_env.step(a, k_substeps=k)
        o2, r2, d2, i2 = b_env.step(a, k_substeps=k)
        assert np.array_equal(d1, d2) and np.array_equal(i1['cause'], i2['cause'])
        assert np.array_equal(i1['final_steps'], i2['final_steps']) and np.array_equal(a_env.steps, b_env.steps)
        assert np.array_equal(a_env.episode, b_env.episode) and np.array_equal(a_env.dyn.status, b_env.status)
        assert merr(r2, r1) <= 1e-12 and merr(b_env.x, a_env.dyn.x) <= 1e-12 and merr(o2, o1) <= 1e-6
        assert i2['executed'] == i1['steps_taken'].sum()
        n_done += d1.sum()
    assert n_done > N
