"""The PID-heuristic restatement (oracle/pid_oracle.py) against the reference's own controller
classes, executed from /root/reference/attic/mars/pidcontrollers (numpy-only, loaded by path)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import refshim
from oracle.pid_oracle import LanderHeuristicBatch

PID_PATH = os.path.join(refshim.REFERENCE_ROOT, 'attic', 'mars', 'pidcontrollers', '__init__.py')
pytestmark = pytest.mark.skipif(not os.path.exists(PID_PATH), reason='no /root/reference')


def reference_heuristic():
    spec = importlib.util.spec_from_file_location('ref_pidcontrollers', PID_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class H:        # the controller set and heuristic body of attic/mars/lander3d.py:34-38, 64-87
        def __init__(self):
            self.phi_rate_pid = mod.AngularVelocityPidController()
            self.theta_rate_pid = mod.AngularVelocityPidController()
            self.x_poshold_pid = mod.PositionHoldPidController()
            self.y_poshold_pid = mod.PositionHoldPidController()
            self.descent_pid = mod.DescentPidController()

        def __call__(self, state):
            x, dx, y, dy, z, dz, phi, dphi, theta, dtheta = state
            phi_todo = self.phi_rate_pid.getDemand(dphi) + self.x_poshold_pid.getDemand(y, dy)
            theta_todo = self.theta_rate_pid.getDemand(-dtheta) + self.y_poshold_pid.getDemand(x, dx)
            descent_todo = self.descent_pid.getDemand(z, dz)
            t, r, p = (descent_todo + 1) / 2, phi_todo, theta_todo
            return [t - r - p, t + r + p, t + r - p, t - r + p]
    return H


def test_heuristic_matches_reference_controllers():
    H = reference_heuristic()
    rng = np.random.default_rng(0)
    n, steps = 16, 300
    refs = [H() for _ in range(n)]
    mine = LanderHeuristicBatch(n)
    worst = 0.0
    for t in range(steps):
        obs = (rng.normal(0, 1, (n, 10)) * np.array([3, 1, 3, 1, 5, 2, .3, 1.0, .3, 1.0])).astype(np.float32)
        a = mine.act(obs)
        for i in range(n):
            r = np.array(refs[i](obs[i].astype(np.float64)))
            worst = max(worst, np.max(np.abs(r - a[i]) / np.maximum(np.abs(r), 1)))
    assert worst <= 1e-13, worst
    # the fast-rotation integral reset and the windup clamp were both exercised
    assert np.abs(mine.x_poshold.err_i).max() == 0.2
