"""The oracle restatement (oracle/copter_oracle.py) against the golden vectors recorded from
the unmodified reference (tests/golden/make_golden.py).  CPU only; runs on the GPU box too."""
import json
import os

import numpy as np
import pytest

from oracle.copter_oracle import (DynamicsBatch, EnvBatch, VARIANTS, STATUS_AIRBORNE,
                                  STATUS_CRASHED, STATUS_LANDED, STATUS_LEVELING)


@pytest.fixture(scope='module')
def kat(golden_dir):
    with open(os.path.join(golden_dir, 'kat.json')) as f:
        return json.load(f)


def test_kat1_constant_thrust_descent(kat):
    d = DynamicsBatch(1)
    s = np.zeros((1, 12))
    s[0, 4] = -10
    d.set_state(s)
    for k in range(1, 1001):
        d.set_motors(1.625e-2 * np.ones((1, 4)))
        g = kat['kat1'].get(str(k))
        if g:
            assert d.x[0, 4] == g['z'] and d.x[0, 5] == g['dz']
            assert d.status[0] == g['status'] and d.ticks[0] == g['ticks']
            assert d.get_time()[0] == g['time']
    # the numbers SURVEY.md section 4 quotes, independently of the json
    assert kat['kat1']['742']['z'] == 0.004667474402665579
    assert kat['kat1']['743']['status'] == STATUS_CRASHED and kat['kat1']['743']['ticks'] == 742
    assert kat['kat1']['1000']['ticks'] == 999


def test_kat2_all_axes(kat):
    g = kat['kat2']
    d = DynamicsBatch(1)
    d.set_state(np.array([g['s0']]))
    for k in range(1, 201):
        d.set_motors(np.array([g['motors']]))
        if str(k) in g['after']:
            assert np.array_equal(d.x[0], np.array(g['after'][str(k)]))
    assert g['after']['1'][1] == 0.2014541314404188


def test_kat3_lander_episode(kat):
    g = kat['kat3']
    env = EnvBatch('Lander3D', 1, auto_reset=False)
    env.reset(force=np.array([g['force']], float))
    rewards = []
    for k in range(1, 1001):
        obs, r, done, info = env.step(1.625e-2 * np.ones((1, 4)))
        rewards.append(r[0])
        if k == 1:
            assert np.array_equal(obs[0], np.float32(g['obs1']))
        if done[0]:
            break
    assert k == g['done_step'] == 732
    assert rewards[:3] == g['rewards_first3'] and rewards[-1] == g['last_reward'] == 0.0
    assert float(np.sum(rewards)) == g['ret'] == 176.20884860298776
    assert env.dyn.status[0] == g['final_status'] == STATUS_CRASHED
    assert env.dyn.x[0, 4] == g['final_z'] and env.dyn.x[0, 5] == g['final_dz']


def test_soft_landing_fsm(kat):
    g = kat['soft_landing']
    env = EnvBatch('Lander3D', 1, auto_reset=False)
    env.reset(force=np.zeros((1, 3)))
    env.dyn.set_state(np.array([g['s0']]))
    seq = []
    for tr in g['trace']:
        obs, r, done, info = env.step(g['action'] * np.ones((1, 4)))
        assert env.dyn.status[0] == tr['status'] and r[0] == tr['reward'] and done[0] == tr['done']
        assert np.array_equal(env.dyn.x[0], np.array(tr['state']))
        seq.append(int(env.dyn.status[0]))
    assert seq == [STATUS_AIRBORNE] * 4 + [STATUS_LEVELING, STATUS_LANDED, STATUS_LANDED]
    assert g['trace'][-1]['reward'] == 100.0 and g['trace'][-1]['done']


def test_takeoff_direct_dynamics(kat):
    for mv, trace in kat['takeoff'].items():
        d = DynamicsBatch(1)
        d.set_state(np.zeros((1, 12)))
        assert d.status[0] == STATUS_LANDED
        for k in range(1, 101):
            d.set_motors(float(mv) * np.ones((1, 4)))
            for g in trace:
                if g['call'] == k:
                    assert (d.x[0, 4], d.x[0, 5], d.status[0], d.ticks[0]) == (
                        g['z'], g['dz'], g['status'], g['ticks'])


@pytest.mark.parametrize('variant', list(VARIANTS))
def test_trajectories_with_autoreset(golden_dir, variant):
    g = np.load(os.path.join(golden_dir, 'traj_%s.npz' % variant))
    act = g['action']
    T, N, A = act.shape
    env = EnvBatch(variant, N, seed=int(g['seed']), auto_reset=True)
    env.reset()
    for t in range(T):
        status_before_reset = None
        obs, r, done, info = env.step(act[t].astype(np.float64))
        assert np.array_equal(done, g['done'][t]), (variant, t)
        assert np.array_equal(r, g['reward'][t]), (variant, t)
        assert np.array_equal(np.where(done, info['final_steps'], env.steps), g['steps'][t])
        if t % 10 == 9:
            assert np.array_equal(env.dyn.x, g['state_every10'][t // 10]), (variant, t)
            assert np.array_equal(obs, env.dyn.x[:, list(VARIANTS[variant][1])].astype(np.float32))
    assert np.array_equal(env.dyn.x, g['final_state'])
    assert g['done'].sum() > 10


def test_k_fusion_equals_single_steps_until_done():
    rng = np.random.default_rng(3)
    N, K = 64, 8
    a = (0.0165 * (1 + 0.2 * rng.uniform(-1, 1, (N, 4)))).astype(np.float32).astype(np.float64)
    e1 = EnvBatch('Lander3D', N, seed=5)
    e2 = EnvBatch('Lander3D', N, seed=5)
    e1.reset(); e2.reset()
    for it in range(40):
        obs_k, r_k, d_k, info = e1.step(a, k_substeps=K)
        rs = np.zeros(N); alive = np.ones(N, bool)
        for k in range(K):
            pre = e2.dyn.x.copy(), e2.dyn.status.copy(), e2.steps.copy(), e2.dyn.perturb.copy(), e2.episode.copy()
            obs, r, d, _ = e2.step(a)
            # envs that already finished in this window idle: roll them back
            idle = ~alive
            e2.dyn.x[idle], e2.dyn.status[idle], e2.steps[idle] = pre[0][idle], pre[1][idle], pre[2][idle]
            e2.dyn.perturb[idle], e2.episode[idle] = pre[3][idle], pre[4][idle]
            rs += np.where(alive, r, 0)
            alive &= ~d
        assert np.array_equal(~alive, d_k)
        assert np.allclose(rs, r_k, rtol=0, atol=1e-12)
        assert np.array_equal(e1.dyn.x, e2.dyn.x)
