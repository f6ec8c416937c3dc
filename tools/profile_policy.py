import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_copter_b200 as g
env = g.LanderVec(1 << 23, seed=1, write_obs=False); env.reset()
pol = g.mlp_policy(10, 4, dtype=torch.float32, seed=5)
fused = g.FusedMLPPolicy(env, pol.net)
for _ in range(3): fused()
torch.cuda.synchronize()
