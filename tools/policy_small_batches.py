import os, sys, torch
sys.path.insert(0, '/root/repo')
import gym_copter_b200 as g
for n in (1 << 14, 1 << 17, 1 << 20, 1 << 22):
    env = g.LanderVec(n, seed=1, write_obs=False); env.reset(); env.rollout(100, source='randn')
    pol = g.mlp_policy(10, 4, dtype=torch.float32, seed=5)
    fused = g.FusedMLPPolicy(env, pol.net, out_scale=0.7, out_offset=0.1)
    ro = g.FusedPolicyRollout(env, pol.net, 16, out_scale=0.2 * 0.0166, out_offset=0.0166)
    out = {}
    for tc in ('1', '0'):
        os.environ['COPTER_B200_POLICY_TC'] = tc; os.environ['COPTER_B200_POLICY_ROLLOUT_TC'] = tc
        for name, fn, reps, div in (('policy_us', fused, 200, 1), ('rollout_us_per_step', ro.run, 20, 16)):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            out[name + ('_tc' if tc == '1' else '_mma')] = round(e0.elapsed_time(e1) / reps / div * 1e3, 2)
    print(n, out, flush=True)
