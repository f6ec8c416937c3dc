"""The oracle's alternate dynamics model bits (lift-model thrust, live gyroscopic Omega; world
gravity / air density) against the reference's older dynamics class, executed from
/root/reference/attic/mars/dynamics (numpy-only; loaded with attic/mars on sys.path)."""
import os
import sys

import numpy as np
import pytest

from oracle import refshim
from oracle.copter_oracle import DynamicsBatch, OracleParams

MARS = os.path.join(refshim.REFERENCE_ROOT, 'attic', 'mars')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(MARS, 'dynamics')), reason='no /root/reference')


def load_mars_quad():
    sys.path.insert(0, MARS)
    try:
        for k in [k for k in sys.modules if k == 'dynamics' or k.startswith('dynamics.')]:
            del sys.modules[k]
        from dynamics import MultirotorDynamics
        from dynamics.djiphantom import QuadXAPDynamics
    finally:
        sys.path.remove(MARS)

    class Quad(QuadXAPDynamics):
        def __init__(self, vparams, fps, wparams):
            MultirotorDynamics.__init__(self, vparams, 4, fps, wparams)
    return Quad


VPARAMS = dict(B=5.E-06, D=2.E-06, M=1.380, L=0.350, C_L=0.4, Ix=2, Iy=2, Iz=3, Jr=38E-04, maxrpm=15000)


@pytest.mark.parametrize('world', [dict(G=9.80655, rho=1.225), dict(G=3.721, rho=0.017)])
def test_lift_and_gyro_model_matches_attic_mars_dynamics(world):
    Quad = load_mars_quad()
    rng = np.random.default_rng(3)
    n, steps = 12, 400
    p = OracleParams(G=world['G'], rho=world['rho'], lift_coefficient=0.4, dynamics_model=3)
    mine = DynamicsBatch(n, p)
    s0 = rng.normal(0, 1, (n, 12)) * np.array([2, 1, 2, 1, 3, 1, .2, .3, .2, .3, .3, .2])
    s0[:, 4] = -np.abs(s0[:, 4]) - 1.0
    refs = [Quad(VPARAMS, 100, world) for _ in range(n)]
    for i, r in enumerate(refs):
        r.setState(s0[i])
    mine.set_state(s0)
    # commands around the hover point of this vehicle/world (lift per motor = M G / 4)
    S = .05 * VPARAMS['L'] * 4
    hover_w = np.sqrt(VPARAMS['M'] * world['G'] / 4 / (0.5 * world['rho'] * S * 0.4)) / (VPARAMS['L'] / 2)
    hover = min(0.9, hover_w / (VPARAMS['maxrpm'] * np.pi / 30))
    for t in range(steps):
        m = np.clip(hover * (1 + 0.3 * rng.uniform(-1, 1, (n, 4))), 0, 1)
        if t % 50 == 0:
            f = rng.uniform(-3, 3, (n, 6))
            mine.set_perturb(f)
            for i, r in enumerate(refs):
                r.perturb(f[i])
        mine.set_motors(m)
        for i, r in enumerate(refs):
            r.setMotors(m[i])
            r.update()
            assert np.array_equal(np.array(r.getState()), mine.x[i]), (t, i)
            assert r.getStatus() == mine.status[i]
    assert np.abs(mine.x[:, 7]).max() > 1e-3        # the torques / gyroscopic terms were exercised
