"""
Trajectory export in the format of the reference's `lander.py --save`
(/root/reference lander.py:33-38,48-54): header `t,m1,m2,m3,m4,<STATE_NAMES>`, one `%f` row
per step with t = dt * step index, the four motor commands as passed to step(), and the
observation returned by that step.  `utils/copter-plot.py:36-61` reads this layout.
"""

import numpy as np


class CsvTrajectoryWriter:

    def __init__(self, path, env, env_index=0):
        self.env, self.index, self.steps = env, int(env_index), 0
        self.dt = 1. / env.FRAMES_PER_SECOND
        self.f = open(path, 'w')
        self.f.write('t,' + ','.join('m%d' % k for k in range(1, 5)))
        self.f.write(',' + ','.join(env.STATE_NAMES) + '\n')

    def write(self, action, obs):
        """`action`, `obs`: what was passed to / returned by step() (batched or single)."""
        a = np.asarray(action.detach().cpu() if hasattr(action, 'detach') else action, dtype=np.float64)
        o = np.asarray(obs.detach().cpu() if hasattr(obs, 'detach') else obs, dtype=np.float64)
        a = a[self.index] if a.ndim == 2 else a
        o = o[self.index] if o.ndim == 2 else o
        m = np.resize(a, 4) if a.size in (1, 4) else np.array([a[0], a[1], a[1], a[0]])
        self.f.write('%f' % (self.dt * self.steps))
        self.f.write((',%f' * 4) % tuple(m))
        self.f.write(((',%f' * len(o)) + '\n') % tuple(o))
        self.steps += 1

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
