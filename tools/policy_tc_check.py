"""Developer tool: the tcgen05 policy kernel against the warp-MMA kernel and the PyTorch fp32 network,
and their timings, then the fused policy + step rollout.
   COPTER_B200_POLICY_TC=1|0 COPTER_B200_POLICY_ROLLOUT_TC=1|0 python tools/policy_tc_check.py [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_copter_b200 as g  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
for variant in ('Lander3D', 'Hover3D', 'Lander2D', 'Lander1D'):
    m = n if variant == 'Lander3D' else 100003
    env = g.CopterVecEnv(variant, m, seed=1)
    env.reset()
    env.rollout(300, source='randn')
    pol = g.mlp_policy(env.obs_size, env.action_size, dtype=torch.float32, seed=5)
    for p in pol.net.parameters():
        p.data.mul_(3.0)
    fused = g.FusedMLPPolicy(env, pol.net, out_scale=0.7, out_offset=0.1)
    got = fused().clone()
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = 0.1 + 0.7 * pol.net(env.obs)
    err = (got - ref).abs()
    print('%-9s tc=%s n=%d max err %.4f mean err %.5f' % (variant, os.environ.get('COPTER_B200_POLICY_TC', 'default'), m, err.max().item(), err.mean().item()), flush=True)
    if variant == 'Lander3D':
        for _ in range(5):
            fused()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50):
            fused()
        e1.record(); torch.cuda.synchronize()
        print('  policy kernel: %.4f ms per launch (%d envs)' % (e0.elapsed_time(e1) / 50, m), flush=True)
        T = 16
        ro = g.FusedPolicyRollout(env, pol.net, T, out_scale=0.2 * 0.0166, out_offset=0.0166)
        for _ in range(2):
            ro.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10):
            ro.run()
        e1.record(); torch.cuda.synchronize()
        print('  fused policy + step rollout (tc=%s): %.4f ms per env-step (%d envs, horizon %d)' % (os.environ.get('COPTER_B200_POLICY_ROLLOUT_TC', 'default'), e0.elapsed_time(e1) / (10 * T), m, T), flush=True)
        del ro
    del env, fused
