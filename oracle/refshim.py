"""
TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (simondlevy/gym-copter)
from /root/reference, used to pin the oracle restatement and to generate the golden
vectors under tests/golden/.  Nothing in the product package imports this module.

`gymnasium` is not installed in this image (and there is no network), so the reference's
`import gymnasium as gym` (gym_copter/envs/task.py:15-17, gym_copter/__init__.py:7) is
satisfied by a ~30-line in-memory stand-in that provides only the names the reference
touches (Env, spaces.Box, utils.EzPickle, utils.seeding.np_random, registration.register).
The stand-in carries no arithmetic: every number the reference produces comes from its own
gym_copter/dynamics/__init__.py, gym_copter/envs/task.py and gym_copter/envs/lander.py.

Variants that exist only in the reference's attic/ are reconstituted as thin subclasses of
the LIVE `_Task` / `Lander`, overriding exactly the three hooks the attic files define:
  Lander2D : attic/gym_copter/envs/lander2d.py:43-51  (obs (y,dy,z,dz,phi,dphi); motors [m0,m1,m1,m0])
  Lander1D : attic/gym_copter/envs/lander1d.py:43-49  (obs (z,dz); motors [m0]*4)
  Hover3D  : attic/gym_copter/envs/hover.py:18-21 + hover3d.py:32-37 (reward == 1, obs all 12)
  Hover2D  : attic/gym_copter/envs/hover2d.py:44-50   (obs state[2:8])
  Hover1D  : attic/gym_copter/envs/hover1d.py:44-50   (obs state[4:6])
  Takeoff  : attic/gym_copter/envs/takeoff.py:18-91   (its own gym.Env, NOT a _Task: setMotors with the
             unclipped action whatever the status, reward = change of -|altitude - 5|, never done).  The
             module it imports (gym_copter.dynamics.djiphantom, setMotors + update) no longer exists; the
             class below is that file's reset()/step() line for line over the LIVE `Dynamics`, whose
             setMotors contains what update() used to do.
"""

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('GYM_COPTER_REFERENCE', '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'gym_copter', 'envs', 'task.py'))


def _install_gymnasium_shim():
    if 'gymnasium' in sys.modules:
        return
    gym = types.ModuleType('gymnasium')

    class Env:
        metadata = {}

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    class EzPickle:
        def __init__(self, *a, **k):
            pass

    spaces = types.ModuleType('gymnasium.spaces')
    spaces.Box = Box
    utils = types.ModuleType('gymnasium.utils')
    seeding = types.ModuleType('gymnasium.utils.seeding')
    seeding.np_random = lambda seed=None: (np.random.default_rng(seed), seed)
    utils.EzPickle = EzPickle
    utils.seeding = seeding
    envs = types.ModuleType('gymnasium.envs')
    registration = types.ModuleType('gymnasium.envs.registration')
    registration.registry = {}
    registration.register = lambda id, **kw: registration.registry.__setitem__(id, kw)
    envs.registration = registration
    gym.Env, gym.spaces, gym.utils, gym.envs = Env, spaces, utils, envs
    gym.__shim__ = True
    sys.modules.update({
        'gymnasium': gym, 'gymnasium.spaces': spaces, 'gymnasium.utils': utils,
        'gymnasium.utils.seeding': seeding, 'gymnasium.envs': envs,
        'gymnasium.envs.registration': registration})


_cache = {}


def load_reference():
    """Returns a namespace with the reference's live classes and the attic-defined variants."""
    if 'ns' in _cache:
        return _cache['ns']
    if not reference_available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    _install_gymnasium_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from gym_copter.dynamics import Dynamics
    from gym_copter.dynamics.vehicles.dji_phantom import vehicle_params
    from gym_copter.envs.task import _Task
    from gym_copter.envs.lander import Lander

    class Lander2D(Lander):
        def __init__(self):
            _Task.__init__(self, 6, 2)

        def _get_state(self, s):
            return (s['y'], s['dy'], s['z'], s['dz'], s['phi'], s['dphi'])

        def _get_motors(self, m):
            return [m[0], m[1], m[1], m[0]]

    class Lander1D(Lander):
        def __init__(self):
            _Task.__init__(self, 2, 1)

        def _get_state(self, s):
            return (s['z'], s['dz'])

        def _get_motors(self, m):
            return [m[0]] * 4

    class Hover3D(_Task):
        def __init__(self):
            _Task.__init__(self, 12, 4)

        def reset(self, seed=None, options=None):
            return _Task._reset(self, seed, options)

        def _get_reward(self, status, state, d, x, y):
            return 1

        def _get_state(self, s):
            return [s[k] for k in ('x', 'dx', 'y', 'dy', 'z', 'dz',
                                   'phi', 'dphi', 'theta', 'dtheta', 'psi', 'dpsi')]

        def _get_motors(self, m):
            return m

    class Hover2D(Hover3D):
        def __init__(self):
            _Task.__init__(self, 6, 2)

        def _get_state(self, s):
            return Hover3D._get_state(self, s)[2:8]

        def _get_motors(self, m):
            return [m[0], m[1], m[1], m[0]]

    class Hover1D(Hover3D):
        def __init__(self):
            _Task.__init__(self, 2, 1)

        def _get_state(self, s):
            return Hover3D._get_state(self, s)[4:6]

        def _get_motors(self, m):
            return [m[0]] * 4

    class Takeoff:
        TARGET_ALTITUDE = 5                                             # takeoff.py:20
        FRAMES_PER_SECOND = 50                                          # takeoff.py:21

        def reset(self):                                                # takeoff.py:45-55
            self.prev_shaping = None
            self.dynamics = Dynamics(vehicle_params, self.FRAMES_PER_SECOND)
            self.dynamics.setState(np.zeros(12))
            return self.step(np.array([0, 0, 0, 0]))[0]

        def step(self, action):                                         # takeoff.py:57-88
            d = self.dynamics
            d.setMotors(action)                                         # (+ d.update() in the attic file)
            s = d.getState()
            state = np.array([s[k] for k in ('x', 'dx', 'y', 'dy', 'z', 'dz', 'phi', 'dphi', 'theta', 'dtheta')])
            altitude = -s['z']
            shaping = -abs(altitude - self.TARGET_ALTITUDE)
            reward = (shaping - self.prev_shaping) if (self.prev_shaping is not None) else 0
            self.prev_shaping = shaping
            return np.array(state, dtype=np.float32), reward, False, {}

    ns = types.SimpleNamespace(
        Dynamics=Dynamics, vehicle_params=vehicle_params, _Task=_Task, Takeoff=Takeoff,
        Lander=Lander, Lander3D=Lander, Lander2D=Lander2D, Lander1D=Lander1D,
        Hover3D=Hover3D, Hover2D=Hover2D, Hover1D=Hover1D)
    _cache['ns'] = ns
    return ns


def reference_reset_with_force(env, force_xyz):
    """
    reset() the reference env, then overwrite the (unseedable, envs/task.py:147,199-202)
    random reset force with a known one.  Nothing has consumed the perturbation yet because
    the priming step inside _reset skips setMotors (envs/task.py:93,197).
    """
    obs, info = env.reset()
    f = np.array([force_xyz[0], force_xyz[1], force_xyz[2], 0.0, 0.0, 0.0], dtype=np.float64)
    env.dynamics.perturb(f)
    return obs, info
