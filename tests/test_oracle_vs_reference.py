"""Live differential test of the oracle against the UNMODIFIED reference, executed from
/root/reference through oracle/refshim.py.  Skipped where the reference tree is absent (the
GPU box); tests/test_oracle_golden.py covers that case with recorded vectors."""
import numpy as np
import pytest

from oracle import refshim
from oracle.copter_oracle import EnvBatch, VARIANTS

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason='no /root/reference')


@pytest.mark.parametrize('variant', list(VARIANTS))
def test_bit_exact_against_reference(variant):
    ref = refshim.load_reference()
    rng = np.random.default_rng(sum(map(ord, variant)))
    N, T = 16, 1001
    A = VARIANTS[variant][2]
    forces = rng.uniform(-30, 30, (N, 3))
    actions = np.empty((T, N, A))
    for i in range(N):
        actions[:, i] = [1.625e-2 * np.ones((T, A)), 1.625e-2 * rng.standard_normal((T, A)),
                         0.016560178 * (1 + 0.1 * rng.uniform(-1, 1, (T, A))),
                         rng.uniform(-1, 1, (T, A))][i % 4]
    env = EnvBatch(variant, N, auto_reset=False)
    env.reset(force=forces)
    refs = []
    for i in range(N):
        e = getattr(ref, variant)()
        refshim.reference_reset_with_force(e, forces[i])
        refs.append(e)
    alive = np.ones(N, bool)
    for t in range(T):
        obs, r, done, info = env.step(actions[t])
        for i in np.nonzero(alive)[0]:
            o_ref, r_ref, d_ref, _, _ = refs[i].step(actions[t, i])
            assert d_ref == done[i] and r_ref == r[i]
            assert np.array_equal(o_ref, obs[i]) and o_ref.dtype == np.float32
            assert np.array_equal(refs[i].dynamics._x, env.dyn.x[i])
            assert refs[i].dynamics.getStatus() == env.dyn.status[i]
            assert refs[i].steps == env.steps[i]
            assert refs[i].dynamics._ticks == env.dyn.ticks[i]
            alive[i] = not done[i]
    assert not alive.any()      # every episode ends by the 1000-step limit at the latest


def test_stepping_past_done_matches_reference():
    """The reference has no auto-reset; stepping a finished env keeps going."""
    ref = refshim.load_reference()
    e = ref.Lander()
    refshim.reference_reset_with_force(e, [1.0, 2.0, 3.0])
    env = EnvBatch('Lander3D', 1, auto_reset=False)
    env.reset(force=np.array([[1.0, 2.0, 3.0]]))
    for t in range(800):
        a = 1.625e-2 * np.ones(4)
        o_ref, r_ref, d_ref, _, _ = e.step(a)
        obs, r, done, _ = env.step(a[None])
        assert (r_ref, d_ref) == (r[0], done[0]) and np.array_equal(o_ref, obs[0])


def test_takeoff_variant_bit_exact_against_the_attic_env():
    """The Takeoff variant against attic/gym_copter/envs/takeoff.py restated over the live Dynamics
    (oracle/refshim.py): unclipped commands of either sign, take-off from the ground, re-landing,
    crashes; the attic env never ends an episode, so the oracle's step limit is put out of reach."""
    from oracle.copter_oracle import OracleParams
    ref = refshim.load_reference()
    rng = np.random.default_rng(5)
    N, T = 12, 900
    actions = np.empty((T, N, 4))
    for i in range(N):
        base = rng.uniform(0.012, 0.022, (T, 1)) * (1 + 0.05 * rng.uniform(-1, 1, (T, 4)))
        base[rng.random(T) < 0.02] = 0.0                       # motors cut: fall back to the ground
        if i % 4 == 1:
            base[300:420] = 0.0                                # a long cut from altitude: hits the ground hard (CRASHED)
        if i % 3 == 0:
            base = -base                                       # no clip (takeoff.py:64): the sign does not matter
        actions[:, i] = base
    p = OracleParams(initial_altitude=0.0, initial_random_force=0.0, fps=50, max_steps=10 ** 6)
    env = EnvBatch('Takeoff', N, p, auto_reset=False)
    obs0 = env.reset()
    refs = [ref.Takeoff() for _ in range(N)]
    for i, e in enumerate(refs):
        assert np.array_equal(e.reset(), obs0[i])
    seen = set()
    for t in range(T):
        obs, r, done, info = env.step(actions[t])
        assert not done.any()
        for i, e in enumerate(refs):
            o_ref, r_ref, d_ref, _ = e.step(actions[t, i])
            assert r_ref == r[i] and d_ref is False
            assert np.array_equal(o_ref, obs[i]) and np.array_equal(e.dynamics._x, env.dyn.x[i])
            assert e.dynamics.getStatus() == env.dyn.status[i]
            seen.add(int(env.dyn.status[i]))
    assert seen == {0, 1, 2, 3} and (env.dyn.x[:, 4] < -1).any()      # every flight status was visited; some fly high
