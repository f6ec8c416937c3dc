#!/usr/bin/env python3
"""
Developer tool: the host-array step (`CopterVecEnv.step_host`, the bench's `e2e`) over chunk sizes
and stream counts, wall clock per step, 2^24 Lander3D envs.

    gpurun -- python tools/sweep_e2e.py
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_copter_b200 as g


def main():
    n = 1 << 24
    for streams in (2, 3, 4, 8):
        for chunk_log2 in (18, 19, 20, 21, 22):
            env = g.LanderVec(n, seed=0)
            env.reset()
            h = env.host_buffers()
            h['action'][:] = 0.0166
            for _ in range(2):
                env.step_host(None, chunk_envs=1 << chunk_log2, n_streams=streams)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 8
            for _ in range(reps):
                env.step_host(None, chunk_envs=1 << chunk_log2, n_streams=streams)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / reps * 1e3
            print(json.dumps({'streams': streams, 'chunk_envs': 1 << chunk_log2, 'ms_per_step': round(ms, 3),
                              'env_steps_per_s': n / ms * 1e3}), flush=True)
            env.close()
            del env, h
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
