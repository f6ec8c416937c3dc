"""The per-object Python port timed by bench.py's reference arm, checked against the golden
vectors (everywhere) and the executed reference (where /root/reference exists)."""
import json
import os
import time

import numpy as np
import pytest

from oracle import refshim
from oracle.scalar_port import ScalarLander, run_stream


def test_kat3(golden_dir):
    g = json.load(open(os.path.join(golden_dir, 'kat.json')))['kat3']
    env = ScalarLander()
    env.reset(force=g['force'])
    rewards = []
    for k in range(1, 1001):
        obs, r, done, _, _ = env.step(1.625e-2 * np.ones(4))
        rewards.append(r)
        if k == 1:
            assert np.array_equal(obs, np.float32(g['obs1']))
        if done:
            break
    assert k == g['done_step'] and rewards[:3] == g['rewards_first3']
    assert float(np.sum(rewards)) == g['ret']


@pytest.mark.skipif(not refshim.reference_available(), reason='no /root/reference')
def test_matches_reference_and_its_speed():
    ref = refshim.load_reference()
    rng = np.random.default_rng(0)
    for ep in range(6):
        f = rng.uniform(-30, 30, 3)
        a_env, b_env = ref.Lander(), ScalarLander()
        refshim.reference_reset_with_force(a_env, f)
        b_env.reset(force=f)
        for t in range(1000):
            a = [1.625e-2 * np.ones(4), 1.625e-2 * rng.standard_normal(4), rng.uniform(-1, 1, 4)][ep % 3]
            o1, r1, d1, _, _ = a_env.step(a)
            o2, r2, d2, _, _ = b_env.step(a)
            assert np.array_equal(o1, o2) and r1 == r2 and d1 == d2
            if d1:
                break
    # throughput of the port is within 2x of the reference's own step (same execution style)
    env = ref.Lander()
    env.reset()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < 1.0:
        _, _, d, _, _ = env.step(1.625e-2 * rng.standard_normal(4))
        n += 1
        if d:
            env.reset()
    ref_rate = n / (time.perf_counter() - t0)
    m, el = run_stream('randn', 1.0)
    assert 0.5 < (m / el) / ref_rate < 2.5, (m / el, ref_rate)
