"""
Batched facade of the reference's `Dynamics` object
(/root/reference: gym_copter/dynamics/__init__.py:33-229): N vehicles driven directly by
motor commands, all four flight statuses reachable (take-off from the ground included).
The method names are the reference's; every method takes / returns a leading N dimension.
Arithmetic: copter_dynamics_{f32,f64} in libcopter_b200.so, nothing on the CPU.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CopterError

STATE_KEYS = ('x', 'dx', 'y', 'dy', 'z', 'dz', 'phi', 'dphi', 'theta', 'dtheta', 'psi', 'dpsi')


class Dynamics:

    (STATE_X, STATE_X_DOT, STATE_Y, STATE_Y_DOT, STATE_Z, STATE_Z_DOT, STATE_PHI, STATE_PHI_DOT,
     STATE_THETA, STATE_THETA_DOT, STATE_PSI, STATE_PSI_DOT) = range(12)
    STATUS_CRASHED, STATUS_LANDED, STATUS_LEVELING, STATUS_AIRBORNE = range(4)

    def __init__(self, params=None, framesPerSecond=100, num=1, dtype=torch.float64, device=None):
        """`params`: the reference's vehicle dict (keys B D M L Ix Iy Iz Jr maxrpm) or None
        for the DJI Phantom (dynamics/vehicles/dji_phantom.py:9-26)."""
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise CopterError('gym_copter_b200 needs a CUDA device (there is no CPU fallback)')
        if dtype not in (torch.float32, torch.float64):
            raise ValueError('dtype must be torch.float32 or torch.float64')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.num, self.dtype = int(num), dtype
        self.params = _lib.default_params(fps=float(framesPerSecond), **(dict(params) if params else {}))
        self._dt = 1. / framesPerSecond
        V = 4 if dtype == torch.float32 else 2
        self._planes = torch.zeros((12 // V, self.num, V), dtype=dtype, device=self.device)
        self._status = torch.full((self.num,), self.STATUS_LANDED, dtype=torch.uint8, device=self.device)
        self._ticks = torch.zeros(self.num, dtype=torch.int32, device=self.device)
        self._perturb = torch.zeros((self.num, 6), dtype=dtype, device=self.device)
        self.launches = 0

    def _t(self, v, cols):
        t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
        t = t.to(device=self.device, dtype=self.dtype)
        return t.expand(self.num, cols).contiguous() if t.dim() == 1 else t.reshape(self.num, cols).contiguous()

    def setMotors(self, motorvals):
        """dynamics/__init__.py:114-197 for every vehicle. motorvals: [N,4] (or [4], broadcast)."""
        m = self._t(motorvals, 4)
        fn = self._lib.copter_dynamics_f32 if self.dtype == torch.float32 else self._lib.copter_dynamics_f64
        with torch.cuda.device(self.device):
            _lib.check(fn(C.byref(self.params), self._planes.data_ptr(), self._status.data_ptr(),
                          self._ticks.data_ptr(), self._perturb.data_ptr(), m.data_ptr(), self.num,
                          C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), 'copter_dynamics')
        self.launches += 1

    def getState(self):
        """dict of [N] tensors keyed like the reference's (dynamics/__init__.py:199-207)."""
        s = self.state
        return {k: s[:, j] for j, k in enumerate(STATE_KEYS)}

    @property
    def state(self):
        return self._planes.permute(1, 0, 2).reshape(self.num, 12)

    def setState(self, state):
        """dynamics/__init__.py:210-217: AIRBORNE iff z < 0, else LANDED."""
        s = self._t(state, 12)
        V = self._planes.shape[2]
        self._planes.copy_(s.reshape(self.num, 12 // V, V).permute(1, 0, 2))
        self._status.copy_(torch.where(s[:, 4] < 0, self.STATUS_AIRBORNE, self.STATUS_LANDED).to(torch.uint8))

    def getTime(self):
        return self._ticks.to(torch.float64) * self._dt              # :219-221

    def getStatus(self):
        return self._status                                          # :223-225

    def perturb(self, force):
        """dynamics/__init__.py:227-229: force [N,6] (or [6]) in newtons / newton-metres."""
        self._perturb.copy_(self._t(force, 6) / self.params.M)
