#!/usr/bin/env python3
"""
Developer tool, run on the GPU box under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py

Exercises every kernel on ragged sizes (partial warps / partial tiles), every variant, both
precisions, statistics, final observations, injected forces, K-fusion, the fused rollouts with
every action source (drawn and PID heuristics), the
policy kernels and the host-array pipeline, so that out-of-bounds accesses and shared-memory hazards in the
obs staging tiles would be reported.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_copter_b200 as g      # noqa: E402

rng = np.random.default_rng(0)
for variant in ('Lander3D', 'Lander2D', 'Lander1D', 'Hover3D', 'Hover2D', 'Hover1D', 'Takeoff'):
    for dtype in (torch.float32, torch.float64):
        for n in (1, 33, 129, 300):
            env = g.CopterVecEnv(variant, n, dtype=dtype, seed=1, k_substeps=3, track_returns=True, keep_final_obs=True,
                                 report_cause=True, wide_counters=(n == 129))
            env.reset(force=rng.uniform(-30, 30, (n, 3)))
            for t in range(4):
                a = rng.uniform(-1, 1, (n, env.action_size)).astype(np.float32)
                env.step(a)
            env.rollout(5, source='randn', record_rewards=True, record_dones=True, record_actions=True)
            env.rollout(3, source='uniform')
            env.rollout(4, source='pid', scale=2e-3, offset=0.0149, record_actions=True)            # 3-D / planar landing heuristics
            if variant not in ('Lander3D', 'Takeoff'):                                             # hover heuristics ([N,24] controller memories)
                env.rollout(4, source='pid_hover', scale=0.0331, record_actions=True)
            env.stats()
for variant in ('Lander3D', 'Lander2D', 'Lander1D', 'Hover3D'):
    for n in (1, 33, 129, 300, 128 * 148 * 4 * 2 + 5):        # the last: more tiles than resident CTAs of the tcgen05 policy kernel
        if n > 1000 and variant != 'Lander3D':
            continue
        env = g.CopterVecEnv(variant, n, seed=4, track_returns=True)
        env.reset()
        pol = g.mlp_policy(env.obs_size, env.action_size, dtype=torch.float32, seed=n % 97)
        for tc in ('1', '0'):                                 # tcgen05 / TMEM kernel, then the warp-MMA kernel
            os.environ['COPTER_B200_POLICY_TC'] = tc
            g.FusedMLPPolicy(env, pol.net, out_scale=0.02, out_offset=0.0166)()
        # (the large size: more tiles than resident CTAs, so the kernels' cluster-launch-control requests succeed)
        ro = g.FusedPolicyRollout(env, pol.net, 6 if n < 1000 else 2, out_scale=0.02, out_offset=0.0166, store_obs=True, store_actions=True)
        for tc in ('1', '0'):                                 # fused rollout on the tcgen05 / TMEM kernel, then on the warp-MMA kernel
            os.environ['COPTER_B200_POLICY_ROLLOUT_TC'] = tc
            ro.run()
            ro.run()
env = g.LanderVec(1000, seed=2, track_stats=True)
env.reset()
for t in range(3):
    env.step_host(rng.uniform(-1, 1, (1000, 4)).astype(np.float32), chunk_envs=256, n_streams=3)
env.close()
d = g.Dynamics(num=77)
d.setState(np.zeros(12))
d.perturb(np.ones(6))
for t in range(5):
    d.setMotors(0.02 * np.ones(4))
torch.cuda.synchronize()
print('sanitize workload finished')
