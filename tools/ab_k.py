"""Developer tool: ms per launch of the fp32 Lander3D step kernel at K = 1, 2, 4, 8, 16 on a DESYNCHRONISED batch
(997 device-side steps on the --random stream first, as bench.py does), for A/B builds:
    COPTER_B200_LIB=tools/variants/lib_X.so python tools/ab_k.py [envs]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_copter_b200 as g  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
acts = [1.625e-2 * torch.randn((n, 4), device='cuda') for _ in range(4)]
out = {}
for k in (1, 2, 4, 8, 16):
    env = g.LanderVec(n, seed=1, k_substeps=k, track_stats=False)
    env.reset()
    env.rollout(997, source='randn')
    for i in range(10):
        env.step(acts[i % 4])
    reps = []
    steps = 100 if k <= 4 else 40
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(steps):
            env.step(acts[i % 4])
        e1.record(); torch.cuda.synchronize()
        reps.append(e0.elapsed_time(e1) / steps)
    out['k%d_ms' % k] = round(sorted(reps)[1], 4)
    del env
print(json.dumps(out))
