"""
TEST INFRASTRUCTURE ONLY -- numpy restatement, batched over envs, of the reference's PID
landing heuristic: attic/mars/pidcontrollers/__init__.py:12-146 (controllers) and
attic/mars/lander3d.py:64-87 (Lander3D.heuristic + quad-X mixer).  Pinned by executing the
reference's controller classes themselves (the module is numpy-only and loads by file path)
in tests/test_pid_oracle.py.  HoverHeuristicBatch is the hover demo's controller set:
attic/mars/hover3d.py:33-38,65-92 with the altitude-hold controller of attic/mars/hover.py:23
(pidcontrollers/__init__.py:70-99).
"""
import numpy as np


class PidBatch:
    """N copies of _PidController (pidcontrollers/__init__.py:12-67)."""

    def __init__(self, n, kp, ki, kd, windup=0.2, dtype=np.float64):
        self.kp, self.ki, self.kd, self.windup = kp, ki, kd, windup
        self.err_i, self.last, self.d1, self.d2 = (np.zeros(n, dtype) for _ in range(4))

    def compute(self, target, actual):
        error = target - actual
        out = error * self.kp
        if self.ki > 0:
            self.err_i = np.clip(self.err_i + error, -self.windup, self.windup)
            out = out + self.err_i * self.ki
        if self.kd > 0:
            de = error - self.last
            out = out + (self.d1 + self.d2 + de) * self.kd
            self.d2, self.d1, self.last = self.d1, de, error
        return out

    def reset_where(self, w):
        self.err_i = np.where(w, 0, self.err_i)
        self.last = np.where(w, 0, self.last)


class LanderHeuristicBatch:
    """Lander3D.heuristic (attic/mars/lander3d.py:34-38, 64-87) for N envs."""

    def __init__(self, n, dtype=np.float64, scale=1.0, offset=0.0, descent_kp=1.15, descent_kd=1.33):
        T = np.dtype(dtype).type
        self.T, self.scale, self.offset = T, T(scale), T(offset)
        self.descent_kp, self.descent_kd = T(descent_kp), T(descent_kd)          # DescentPidController (:110-121)
        self.phi_rate = PidBatch(n, T(1.0), T(0), T(1.0), T(6), dtype)          # AngularVelocityPidController (:124-135)
        self.theta_rate = PidBatch(n, T(1.0), T(0), T(1.0), T(6), dtype)
        self.x_poshold = PidBatch(n, T(0.00001), T(0.1), T(4.0), T(0.2), dtype)  # PositionHoldPidController (:102-107)
        self.y_poshold = PidBatch(n, T(0.00001), T(0.1), T(4.0), T(0.2), dtype)
        self.big = T(np.radians(40))

    def _rate(self, pid, rate):
        pid.reset_where(np.abs(rate) > self.big)                                  # :141-143
        return pid.compute(self.T(0), rate)

    def _poshold(self, pid, x, dx):
        return pid.compute((self.T(0) - x) * self.T(1), dx)                       # :80-88 with posPid(1,0,0)

    def act(self, obs):
        """obs: [N,10] float32 observation of the previous step. Returns motors [N,4]."""
        T = self.T
        o = obs.astype(np.float32).astype(self.phi_rate.err_i.dtype)
        phi_todo = self._rate(self.phi_rate, o[:, 7]) + self._poshold(self.x_poshold, o[:, 2], o[:, 3])
        theta_todo = self._rate(self.theta_rate, -o[:, 9]) + self._poshold(self.y_poshold, o[:, 0], o[:, 1])
        descent_todo = o[:, 4] * self.descent_kp + o[:, 5] * self.descent_kd
        t, r, p = (descent_todo + T(1)) / T(2), phi_todo, theta_todo
        mix = np.stack([t - r - p, t + r + p, t + r - p, t - r + p], -1)          # lander3d.py:87
        return self.offset + self.scale * mix


class HoverHeuristicBatch:
    """Hover3D.heuristic (attic/mars/hover3d.py:33-38, 65-92; altpid from hover.py:23) for N envs."""

    def __init__(self, n, dtype=np.float64, scale=1.0, offset=0.0, alt_target=5.0, alt_kp=0.2, alt_ki=3.0, alt_kd=0.0,
                 alt_windup=0.2):
        T = np.dtype(dtype).type
        self.T, self.scale, self.offset = T, T(scale), T(offset)
        self.roll_rate = PidBatch(n, T(1.0), T(0), T(1.0), T(6), dtype)         # AngularVelocityPidController (:124-135)
        self.pitch_rate = PidBatch(n, T(1.0), T(0), T(1.0), T(6), dtype)
        self.yaw_rate = PidBatch(n, T(1.0), T(0), T(1.0), T(6), dtype)
        self.x_poshold = PidBatch(n, T(0.00001), T(0.1), T(4.0), T(0.2), dtype)  # PositionHoldPidController (:102-107)
        self.y_poshold = PidBatch(n, T(0.00001), T(0.1), T(4.0), T(0.2), dtype)
        self.alt = PidBatch(n, T(alt_kp), T(alt_ki), T(alt_kd), T(alt_windup), dtype)   # AltitudeHoldPidController (:91-99)
        self.alt_target = T(alt_target)
        self.big = T(np.radians(40))

    def _rate(self, pid, rate):
        pid.reset_where(np.abs(rate) > self.big)                                  # :141-143
        return pid.compute(self.T(0), rate)

    def act(self, obs):
        """obs: [N,12] float32 observation of the previous step. Returns motors [N,4]."""
        T = self.T
        o = obs.astype(np.float32).astype(self.alt.err_i.dtype)
        roll_todo = self._rate(self.roll_rate, o[:, 7]) + self.x_poshold.compute((T(0) - o[:, 2]) * T(1), o[:, 3])
        pitch_todo = self._rate(self.pitch_rate, -o[:, 9]) + self.y_poshold.compute((T(0) - o[:, 0]) * T(1), o[:, 1])
        yaw_todo = self._rate(self.yaw_rate, -o[:, 11])
        # AltitudeHoldPidController.getDemand(z, dz): set-point controller on (-z, -dz)  (:96-99, 80-88)
        hover_todo = self.alt.compute((self.alt_target - (-o[:, 4])) * T(1), -o[:, 5])
        t, r, p, y = (hover_todo + T(1)) / T(2), roll_todo, pitch_todo, yaw_todo
        mix = np.stack([t - r - p - y, t + r + p - y, t + r - p + y, t - r + p + y], -1)   # hover3d.py:92
        return self.offset + self.scale * mix


class PlanarHeuristicBatch:
    """The 2-D and 1-D heuristic demos (attic/heuristic/lander2d.py:14-24, lander1d.py:14-20,
    hover2d.py:17-31, hover1d.py:14-20) for N envs.  `kind` = 'lander' | 'hover', `dims` = 2 | 1.
    Observation: (y, dy, z, dz, phi, dphi) for 2-D, (z, dz) for 1-D.  No (t+1)/2 and no mixer here:
    the demand itself is the command (2-D: demand -/+ roll correction for the two motor pairs)."""

    def __init__(self, n, kind, dims, dtype=np.float64, scale=1.0, offset=0.0, descent_kp=1.15, descent_kd=1.33,
                 alt_target=5.0):
        T = np.dtype(dtype).type
        self.T, self.kind, self.dims, self.scale, self.offset = T, kind, dims, T(scale), T(offset)
        self.descent_kp, self.descent_kd = T(descent_kp), T(descent_kd)
        self.rate = PidBatch(n, T(1.0), T(0), T(1.0), T(6), dtype)
        self.poshold = PidBatch(n, T(0.00001), T(0.1), T(4.0), T(0.2), dtype)
        self.alt = PidBatch(n, T(0.2), T(3.0), T(0.0), T(0.2), dtype)
        self.alt_target = T(alt_target)
        self.big = T(np.radians(40))

    def act(self, obs):
        T = self.T
        o = obs.astype(np.float32).astype(self.alt.err_i.dtype)
        z, dz = (o[:, 2], o[:, 3]) if self.dims == 2 else (o[:, 0], o[:, 1])
        if self.kind == 'lander':
            demand = z * self.descent_kp + dz * self.descent_kd
        else:
            demand = self.alt.compute((self.alt_target - (-z)) * T(1), -dz)
        if self.dims == 1:
            return self.offset + self.scale * demand[:, None]
        todo = self.poshold.compute((T(0) - o[:, 0]) * T(1), o[:, 1])
        if self.kind == 'hover':
            self.rate.reset_where(np.abs(o[:, 5]) > self.big)
            todo = self.rate.compute(T(0), o[:, 5]) + todo
        return self.offset + self.scale * np.stack([demand - todo, demand + todo], -1)
