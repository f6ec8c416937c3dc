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


def _load_attic_heuristic(name):
    """Executes /root/reference/attic/heuristic/<name>.py unmodified and returns its `heuristic`
    function.  The scripts import `main.demo` (absent from the reference tree: the demo runner) and
    run it at import time, so `main` is a stub whose demo() does nothing; `pidcontrollers` is the
    reference's own attic/mars module."""
    import sys
    import types
    spec = importlib.util.spec_from_file_location('pidcontrollers', PID_PATH)
    pidmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pidmod)
    stub = types.ModuleType('main')
    stub.demo = lambda *a, **k: None
    stub.demo3d = lambda *a, **k: None
    saved = {k: sys.modules.get(k) for k in ('main', 'pidcontrollers')}
    sys.modules['main'], sys.modules['pidcontrollers'] = stub, pidmod
    try:
        path = os.path.join(refshim.REFERENCE_ROOT, 'attic', 'heuristic', name + '.py')
        s = importlib.util.spec_from_file_location('ref_heuristic_' + name, path)
        mod = importlib.util.module_from_spec(s)
        s.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.heuristic, pidmod


@pytest.mark.parametrize('name,kind,dims', [('lander2d', 'lander', 2), ('lander1d', 'lander', 1),
                                            ('hover2d', 'hover', 2), ('hover1d', 'hover', 1)])
def test_planar_heuristics_match_reference_scripts(name, kind, dims):
    from oracle.pid_oracle import PlanarHeuristicBatch
    heuristic, pm = _load_attic_heuristic(name)
    make = {'lander2d': lambda: (pm.PositionHoldPidController(), pm.DescentPidController()),
            'lander1d': lambda: (pm.DescentPidController(),),
            'hover2d': lambda: (pm.AngularVelocityPidController(), pm.PositionHoldPidController(), pm.AltitudeHoldPidController()),
            'hover1d': lambda: (pm.AltitudeHoldPidController(),)}[name]
    rng = np.random.default_rng(2)
    n, steps = 8, 300
    refs = [make() for _ in range(n)]
    mine = PlanarHeuristicBatch(n, kind, dims)
    worst = 0.0
    for t in range(steps):
        obs = rng.normal(0, 1, (n, 6)) * np.array([3, 1, 5, 2, .3, 1.0])
        if t % 3:
            obs[:, 2:4] = np.array([-5.0, 0.0]) + 0.02 * rng.normal(0, 1, (n, 2))
        obs = (obs if dims == 2 else obs[:, 2:4]).astype(np.float32)
        a = mine.act(obs)
        for i in range(n):
            r = np.array(heuristic(tuple(obs[i].astype(np.float64)), refs[i]))
            worst = max(worst, np.max(np.abs(r - a[i]) / np.maximum(np.abs(r), 1)))
    assert worst <= 1e-13, worst


def test_hover3d_heuristic_script_equals_the_env_method():
    """attic/heuristic/hover.py:19-48 is the same controller as attic/mars/hover3d.py:65-92."""
    import sys
    import types
    saved = {k: sys.modules.get(k) for k in ('gym_copter', 'gym_copter.rendering', 'gym_copter.rendering.threed')}
    for k in saved:
        sys.modules[k] = types.ModuleType(k)
    sys.modules['gym_copter.rendering.threed'].ThreeDHoverRenderer = None
    try:
        heuristic, pm = _load_attic_heuristic('hover')
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    H = reference_hover_heuristic()
    a, b = H(), tuple(c() for c in (pm.AngularVelocityPidController,) * 3 + (pm.PositionHoldPidController,) * 2 + (pm.AltitudeHoldPidController,))
    rng = np.random.default_rng(3)
    for t in range(100):
        s = rng.normal(0, 1, 12) * np.array([3, 1, 3, 1, 5, 2, .3, 1.0, .3, 1.0, .3, 1.0])
        assert np.allclose(a(s), heuristic(s, b), rtol=0, atol=0)
