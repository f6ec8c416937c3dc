"""Developer tool: the launches `ncu -k regex:copter_` captures for profiles/ beyond the headline
kernel -- the fp64 step (config 4), the Lander2D step (config 2), the reset kernel, the direct
Dynamics kernel and the fused rollout kernel on its drawn action sources.  One launch of each is
the last one of its kind, so `-k regex:... --launch-skip` is not needed: every launch is captured
and tools/ncu_summary.py prints them in order."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_copter_b200 as g  # noqa: E402

# config 4: Hover3D fp64, 2^22 envs
env = g.Hover3DVec(1 << 22, dtype=torch.float64, seed=4)
env.reset()                                                        # launch: copter_reset_kernel<double>
a = 0.01656 * (1 + 0.05 * torch.randn((env.num_envs, 4), device='cuda', dtype=torch.float64))
for _ in range(2):
    env.step(a)                                                    # launches: copter_step_kernel<double, HOVER3D>
del env, a
# config 2: Lander2D fp32, 2^20 and 2^24 envs, U(-1,1) commands
for n in (1 << 20, 1 << 24):
    env = g.Lander2DVec(n, seed=2)
    env.reset()
    a = 2 * torch.rand((n, 2), device='cuda') - 1
    for _ in range(2):
        env.step(a)                                                # copter_step_kernel<float, LANDER2D>
    del env, a
# fused rollout kernel, drawn sources, Lander3D fp32 2^22 envs x 16 steps
env = g.LanderVec(1 << 22, seed=3)
env.reset()
for src in ('const', 'randn', 'uniform'):
    env.rollout(16, source=src)                                    # copter_rollout_kernel<float, LANDER3D>
# PID landing heuristic closed on the device
env.rollout(16, source='pid', scale=2e-3, offset=0.0149, pid_gains={'descent_kd': 3.0})
del env
# direct Dynamics facade, 2^20 vehicles
d = g.Dynamics(num=1 << 20)
d.setState(np.zeros(12))
for _ in range(2):
    d.setMotors(0.02 * np.ones(4))                                 # copter_dynamics_kernel<double>
torch.cuda.synchronize()
