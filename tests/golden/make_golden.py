#!/usr/bin/env python3
"""
Generates the golden vectors under tests/golden/ by EXECUTING the unmodified reference
(/root/reference, loaded through oracle/refshim.py).  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box and ships no fixtures of its own, so these files
are what pins the oracle (and through it the CUDA path) there.

Files:
  kat.json            known-answer vectors: KAT-1/2/3 of SURVEY.md section 4, the soft-landing
                      FSM trace, the direct-Dynamics take-off trace, ticks/getTime behaviour.
  traj_<Variant>.npz  12 envs x 1000 steps, four action streams, same-step auto-reset emulated
                      on the reference by reset() + injected force (force = the product's
                      Philox draw for (seed, env, episode); Philox itself is pinned by the
                      Random123 vectors in tests/test_philox.py).
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim                                    # noqa: E402
from oracle.copter_oracle import VARIANTS, reset_force        # noqa: E402

N_ENVS, N_STEPS, SEED = 12, 1000, 20261017
HOVER = 0.016560178212092172


def action_streams(rng, n, t, a):
    """fp32-representable action streams, one kind per env (kind = env index mod 4)."""
    act = np.empty((t, n, a), np.float32)
    for i in range(n):
        kind = i % 4
        if kind == 0:       # lander.py:21,42 constant-thrust heuristic
            act[:, i] = 1.625e-2
        elif kind == 1:     # lander.py:42 --random
            act[:, i] = 1.625e-2 * rng.standard_normal((t, a))
        elif kind == 2:     # hover +-10 %: long episodes
            act[:, i] = HOVER * (1 + 0.1 * rng.uniform(-1, 1, (t, a)))
        else:               # action-space uniform: reset-dominated
            act[:, i] = rng.uniform(-1, 1, (t, a))
    return act


def make_traj(ref, variant):
    rng = np.random.default_rng(SEED + sum(map(ord, variant)))
    a = VARIANTS[variant][2]
    act = action_streams(rng, N_ENVS, N_STEPS, a)
    reward = np.zeros((N_STEPS, N_ENVS))
    done = np.zeros((N_STEPS, N_ENVS), bool)
    status = np.zeros((N_STEPS, N_ENVS), np.int8)     # after the step, before auto-reset
    steps = np.zeros((N_STEPS, N_ENVS), np.int16)     # idem
    state = np.zeros((N_STEPS, N_ENVS, 12))           # after the step AND auto-reset
    for i in range(N_ENVS):
        env = getattr(ref, variant)()
        episode = 0
        f = reset_force(SEED, [i], [episode], 30.0)[0]
        refshim.reference_reset_with_force(env, f)
        for t in range(N_STEPS):
            obs, r, d, trunc, info = env.step(act[t, i].astype(np.float64))
            reward[t, i], done[t, i] = r, d
            status[t, i], steps[t, i] = env.dynamics.getStatus(), env.steps
            if d:
                episode += 1
                f = reset_force(SEED, [i], [episode], 30.0)[0]
                obs, _ = refshim.reference_reset_with_force(env, f)
            state[t, i] = env.dynamics._x
            assert np.array_equal(obs, np.float32([state[t, i][j] for j in VARIANTS[variant][1]]))
    np.savez_compressed(
        os.path.join(HERE, 'traj_%s.npz' % variant), seed=SEED, action=act, reward=reward,
        done=done, status=status, steps=steps, state_every10=state[9::10].copy(),
        final_state=state[-1].copy())
    print(variant, 'episodes finished:', int(done.sum()))


def make_kat(ref):
    D, vp = ref.Dynamics, ref.vehicle_params
    kat = {}

    # KAT-1: free descent under the constant-thrust heuristic, crash at call 743
    d = D(vp, 100)
    s = np.zeros(12)
    s[4] = -10
    d.setState(s)
    rec = {}
    for k in range(1, 1001):
        d.setMotors(1.625e-2 * np.ones(4))
        if k in (1, 100, 500, 742, 743, 744, 1000):
            rec[k] = dict(z=d._x[4], dz=d._x[5], status=int(d.getStatus()),
                          ticks=int(d._ticks), time=d.getTime())
    kat['kat1'] = rec

    # KAT-2: all-axes trajectory
    d = D(vp, 100)
    s0 = [1.0, 0.2, -2.0, -0.1, -8.0, 0.3, 0.05, 0.01, -0.02, 0.02, 0.1, -0.03]
    m = [0.0160, 0.0165, 0.0170, 0.0166]
    d.setState(s0)
    rec = {}
    for k in range(1, 201):
        d.setMotors(np.array(m))
        if k in (1, 200):
            rec[k] = list(d._x)
    kat['kat2'] = dict(s0=s0, motors=m, after=rec)

    # KAT-3: Lander env, injected force, constant action
    env = ref.Lander()
    refshim.reference_reset_with_force(env, [10, -5, 3])
    rewards, obs1 = [], None
    for k in range(1, 1001):
        obs, r, done, _, _ = env.step(1.625e-2 * np.ones(4))
        rewards.append(r)
        if k == 1:
            obs1 = [float(v) for v in obs]
        if done:
            break
    kat['kat3'] = dict(force=[10, -5, 3], obs1=obs1, rewards_first3=rewards[:3], done_step=k,
                       final_status=int(env.dynamics.getStatus()), last_reward=rewards[-1],
                       ret=float(np.sum(rewards)), final_z=env.dynamics._x[4],
                       final_dz=env.dynamics._x[5])

    # soft-landing FSM trace: LEVELING at step 5, LANDED at 6, done + bonus at 7
    env = ref.Lander()
    env.reset()
    env.dynamics.perturb(np.zeros(6))
    s = np.zeros(12)
    s[0], s[2], s[4], s[5], s[6], s[8] = 0.5, -0.3, -0.02, 0.5, 0.1, -0.05
    env.dynamics.setState(s)
    env.prev_shaping = None
    env.step(np.zeros(4), initializing=True)      # prime prev_shaping as _reset does
    env.steps = 1
    trace = []
    for k in range(1, 9):
        obs, r, done, _, _ = env.step(0.01656 * np.ones(4))
        trace.append(dict(step=k, status=int(env.dynamics.getStatus()), reward=r, done=bool(done),
                          state=list(env.dynamics._x)))
        if done:
            break
    kat['soft_landing'] = dict(s0=list(s), action=0.01656, trace=trace)

    # take-off by driving Dynamics directly (config 4's takeoff-style oracle)
    rec = {}
    for mv in (0.02, 0.01):
        d = D(vp, 100)
        d.setState(np.zeros(12))
        tr = []
        for k in range(1, 101):
            d.setMotors(mv * np.ones(4))
            if k in (1, 2, 100):
                tr.append(dict(call=k, z=d._x[4], dz=d._x[5], status=int(d.getStatus()),
                               ticks=int(d._ticks)))
        rec[str(mv)] = tr
    kat['takeoff'] = rec

    with open(os.path.join(HERE, 'kat.json'), 'w') as f:
        json.dump(kat, f, indent=1)
    print('kat.json written')


if __name__ == '__main__':
    ref = refshim.load_reference()
    make_kat(ref)
    for v in VARIANTS:
        make_traj(ref, v)
