"""Developer tool: host-array step (CopterVecEnv.step_host) per shard size with the direct (zero-copy) path forced on
or off -- where does one launch over mapped host memory stop beating the chunked copy pipeline?
    COPTER_B200_DIRECT_MAX_ENVS=<n> python tools/direct_path_sweep.py      (one process per setting: the limit is read once)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_copter_b200 as g  # noqa: E402

out = {}
for n in [int(x) for x in sys.argv[1:]] or (1, 256, 1024, 4096, 16384, 65536, 262144):
    env = g.LanderVec(n, seed=1)
    env.reset()
    a = (1.625e-2 * np.ones((n, 4))).astype(np.float32)
    for _ in range(20):
        env.step_host(a)
    reps = 300 if n <= 65536 else (100 if n <= (1 << 20) else 10)
    t0 = time.perf_counter()
    for _ in range(reps):
        env.step_host(a)
    out[n] = round((time.perf_counter() - t0) / reps * 1e6, 1)
    env.close()
print(os.environ.get('COPTER_B200_DIRECT_MAX_ENVS', 'default'), json.dumps(out))
