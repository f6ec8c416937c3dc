"""Developer tool: the launches `ncu -k regex:policy` captures for profiles/ -- the tcgen05 / TMEM policy
kernel (x2), the warp-MMA policy kernel (x2), then the fused policy + step rollout kernels (tcgen05 x2, warp-MMA x2), 2^23 envs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_copter_b200 as g  # noqa: E402

env = g.LanderVec(1 << 23, seed=1, write_obs=False)
env.reset()
pol = g.mlp_policy(10, 4, dtype=torch.float32, seed=5)
fused = g.FusedMLPPolicy(env, pol.net, out_scale=0.2 * 0.0166, out_offset=0.0166)
for tc in ('1', '0'):
    os.environ['COPTER_B200_POLICY_TC'] = tc
    for _ in range(2):
        fused()
ro = g.FusedPolicyRollout(env, pol.net, 8, out_scale=0.2 * 0.0166, out_offset=0.0166)
for tc in ('1', '0'):                       # the fused policy + step rollout: tcgen05 / TMEM kernel (x2), warp-MMA kernel (x2)
    os.environ['COPTER_B200_POLICY_ROLLOUT_TC'] = tc
    for _ in range(2):
        ro.run()
torch.cuda.synchronize()
