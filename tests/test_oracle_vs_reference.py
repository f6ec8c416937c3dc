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
