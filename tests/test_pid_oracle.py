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


def reference_hover_heuristic():
    spec = importlib.util.spec_from_file_location('ref_pidcontrollers_h', PID_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class H:        # the controller set and heuristic body of attic/mars/hover3d.py:33-38, 65-92 (+ hover.py:23)
        def __init__(self):
            self.roll_rate_pid = mod.AngularVelocityPidController()
            self.pitch_rate_pid = mod.AngularVelocityPidController()
            self.yaw_rate_pid = mod.AngularVelocityPidController()
            self.x_poshold_pid = mod.PositionHoldPidController()
            self.y_poshold_pid = mod.PositionHoldPidController()
            self.altpid = mod.AltitudeHoldPidController()

        def __call__(self, state):
            x, dx, y, dy, z, dz, phi, dphi, theta, dtheta, _, dpsi = state
            roll_todo = self.roll_rate_pid.getDemand(dphi) + self.x_poshold_pid.getDemand(y, dy)
            pitch_todo = self.pitch_rate_pid.getDemand(-dtheta) + self.y_poshold_pid.getDemand(x, dx)
            yaw_todo = self.yaw_rate_pid.getDemand(-dpsi)
            hover_todo = self.altpid.getDemand(z, dz)
            t, r, p, y = (hover_todo + 1) / 2, roll_todo, pitch_todo, yaw_todo
            return [t - r - p - y, t + r + p - y, t + r - p + y, t - r + p + y]
    return H


def test_hover_heuristic_matches_reference_controllers():
    from oracle.pid_oracle import HoverHeuristicBatch
    H = reference_hover_heuristic()
    rng = np.random.default_rng(1)
    n, steps = 16, 300
    refs = [H() for _ in range(n)]
    mine = HoverHeuristicBatch(n)
    worst, unclamped = 0.0, 0
    for t in range(steps):
        obs = rng.normal(0, 1, (n, 12)) * np.array([3, 1, 3, 1, 5, 2, .3, 1.0, .3, 1.0, .3, 1.0])
        if t % 3:                    # near the set-point too, so that the altitude integrator leaves its clamp
            obs[:, 4:6] = np.array([-5.0, 0.0]) + 0.02 * rng.normal(0, 1, (n, 2))
        obs = obs.astype(np.float32)
        a = mine.act(obs)
        unclamped += int((np.abs(mine.alt.err_i) < 0.2).sum())
        for i in range(n):
            r = np.array(refs[i](obs[i].astype(np.float64)))
            worst = max(worst, np.max(np.abs(r - a[i]) / np.maximum(np.abs(r), 1)))
    assert worst <= 1e-13, worst
    assert np.abs(mine.alt.err_i).max() == 0.2 and unclamped > n * steps // 4
