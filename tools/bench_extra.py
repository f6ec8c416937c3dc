#!/usr/bin/env python3
"""
Side benchmarks for the BASELINE.json configs other than the headline (run on the GPU box):

  config2   Lander2D, 2^20 envs, fp32, action-space-uniform and --random streams, 1 GPU
            (launch-bound regime: eager launches vs one CUDA graph of 32 steps)
  rollout   fused multi-step rollouts with on-device const / randn / uniform action sources
  policy    config 5: Lander3D, 2^23 envs/GPU, small tanh MLP policy in the loop, eager vs
            CUDA-graph replay (policy fwd + env step, zero-copy obs/action)
  fp64      config 4: Hover3D fp64 throughput (validation path)

Prints one JSON object per measurement.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gym_copter_b200 as g      # noqa: E402


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def config2():
    n = 1 << 20
    for stream in ('unif', 'randn'):
        env = g.Lander2DVec(n, seed=1)
        env.reset()
        acts = [(2 * torch.rand((n, 2), device='cuda') - 1) if stream == 'unif' else 1.625e-2 * torch.randn((n, 2), device='cuda') for _ in range(8)]
        it = [0]

        def one():
            env.step(acts[it[0] % 8]); it[0] += 1
        for _ in range(50):
            one()
        ms = timed(one, 2000)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for j in range(32):
                env.step(acts[j % 8])
        gr.replay()
        ms_g = timed(gr.replay, 100) / 32
        print(json.dumps({'bench': 'config2 Lander2D 2^20 envs fp32 %s' % stream, 'ms_per_step_eager': ms, 'steps_per_s_eager': n / ms * 1e3,
                          'ms_per_step_graph32': ms_g, 'steps_per_s_graph32': n / ms_g * 1e3,
                          'gbs_graph32': 141 * n / ms_g / 1e6}), flush=True)


def rollout():
    n = 1 << 24
    for src in ('const', 'randn', 'uniform'):
        env = g.LanderVec(n, seed=2, track_stats=True)
        env.reset()
        for T in (16, 64):
            env.rollout(T, source=src)
            before = env.stats()['env_steps']
            ms = timed(lambda: env.rollout(T, source=src), 10)
            steps = (env.stats()['env_steps'] - before) / 10
            print(json.dumps({'bench': 'rollout Lander3D 2^24 envs fp32 source=%s T=%d' % (src, T), 'ms_per_launch': ms,
                              'steps_per_s': steps / ms * 1e3}), flush=True)


def policy():
    n, T = 1 << 23, 8
    for dt in (torch.bfloat16, torch.float32):
        env = g.LanderVec(n, seed=3)
        env.reset()
        pol = g.mlp_policy(10, 4, dtype=dt)
        scaled = lambda obs: 0.0166 * (1 + 0.2 * pol(obs))        # noqa: E731
        for graph in (False, True):
            ro = g.PolicyRollout(env, scaled, T, use_cuda_graph=graph)
            ro.run(); ro.run()
            ms = timed(ro.run, 10)
            print(json.dumps({'bench': 'policy-in-loop Lander3D 2^23 envs, MLP 10-64-64-4 %s, T=%d, %s' % (str(dt).split('.')[-1], T, 'cuda graph' if graph else 'eager'),
                              'ms_per_env_step': ms / T, 'steps_per_s': n * T / ms * 1e3}), flush=True)
        if dt == torch.bfloat16:
            # the same network evaluated by the hand-written policy kernel, zero-copy from the state planes
            env2 = g.LanderVec(n, seed=3, write_obs=False)
            env2.reset()
            pol32 = g.mlp_policy(10, 4, dtype=torch.float32)
            fused = g.FusedMLPPolicy(env2, pol32.net, out_scale=0.2 * 0.0166, out_offset=0.0166)
            ms_pol = timed(fused, 50)
            for graph in (False, True):
                ro = g.PolicyRollout(env2, fused, T, use_cuda_graph=graph, planar=True)
                ro.run(); ro.run()
                ms = timed(ro.run, 10)
                print(json.dumps({'bench': 'policy-in-loop Lander3D 2^23 envs, FUSED MLP kernel 10-64-64-4, T=%d, %s' % (T, 'cuda graph' if graph else 'eager'),
                                  'ms_per_env_step': ms / T, 'steps_per_s': n * T / ms * 1e3, 'policy_kernel_ms': ms_pol}), flush=True)
            # policy + env step fused over the whole horizon: ONE launch, state in registers -- on the tcgen05 / TMEM
            # kernel (the default) and on the warp-MMA kernel
            for tc in ('1', '0'):
                os.environ['COPTER_B200_POLICY_ROLLOUT_TC'] = tc
                for T2, kw in ((8, {}), (32, {}), (32, dict(store_obs=True, store_actions=True)),
                               (32, dict(store_obs=True, store_actions=True, action_std=0.002))):
                    fr = g.FusedPolicyRollout(env2, pol32.net, T2, out_scale=0.2 * 0.0166, out_offset=0.0166, **kw)
                    fr.run(); fr.run()
                    ms = timed(fr.run, 10)
                    print(json.dumps({'bench': 'policy-in-loop Lander3D 2^23 envs, FUSED policy+step rollout kernel (%s), T=%d%s'
                                               % ('tcgen05/TMEM' if tc == '1' else 'warp-MMA', T2, (', obs/action rows recorded' if kw else '') + (', Gaussian exploration noise' if 'action_std' in kw else '')),
                                      'ms_per_env_step': ms / T2, 'steps_per_s': n * T2 / ms * 1e3}), flush=True)
                    del fr
            os.environ.pop('COPTER_B200_POLICY_ROLLOUT_TC', None)
            del env2
        # env-only share of the same loop
        a = torch.full((n, 4), 0.0166, device='cuda')
        ms_env = timed(lambda: env.step(a), 200)
        print(json.dumps({'bench': 'env step alone at 2^23 envs', 'ms_per_env_step': ms_env, 'steps_per_s': n / ms_env * 1e3}), flush=True)
        del env
        torch.cuda.empty_cache()


def zerocopy():
    n = 1 << 24
    acts = [1.625e-2 * torch.randn((n, 4), device='cuda') for _ in range(4)]
    for write_obs in (True, False):
        env = g.LanderVec(n, seed=5, write_obs=write_obs)
        env.reset()
        it = [0]

        def one():
            env.step(acts[it[0] % 4]); it[0] += 1
        for _ in range(20):
            one()
        ms = min(timed(one, 200) for _ in range(3))
        b = 165 if write_obs else 125
        print(json.dumps({'bench': 'Lander3D 2^24 envs fp32 K=1 write_obs=%s' % write_obs, 'ms_per_step': ms,
                          'steps_per_s': n / ms * 1e3, 'bytes_per_env': b, 'gbs': b * n / ms / 1e6}), flush=True)
        del env
        torch.cuda.empty_cache()


def fp64():
    n = 1 << 22
    env = g.Hover3DVec(n, dtype=torch.float64, seed=4)
    env.reset()
    a = 0.01656 * (1 + 0.05 * torch.randn((n, 4), device='cuda', dtype=torch.float64))
    for _ in range(10):
        env.step(a)
    ms = timed(lambda: env.step(a), 100)
    print(json.dumps({'bench': 'config4 Hover3D fp64 2^22 envs', 'ms_per_step': ms, 'steps_per_s': n / ms * 1e3, 'gbs': 289 * n / ms / 1e6}), flush=True)


def single():
    """config 1: the single-env facade (`make('gym_copter:Lander-v0')`) in lander.py's loop --
    latency per step() call (one env cannot use the GPU's width; this is the drop-in call shape)."""
    import time
    import numpy as np
    for dt in (torch.float64, torch.float32):
        env = g.make('gym_copter:Lander-v0', dtype=dt)
        env.reset()
        a = 1.625e-2 * np.ones(4)
        n, steps = 0, 0
        for _ in range(200):
            env.step(a)
        t0 = time.perf_counter()
        while steps < 5000:
            obs, r, done, _, _ = env.step(a)
            steps += 1
            if done:
                env.reset(); n += 1
        el = time.perf_counter() - t0
        print(json.dumps({'bench': 'config1 single-env facade Lander-v0 %s, constant thrust, reset on done' % str(dt).split('.')[-1],
                          'us_per_step': el / steps * 1e6, 'steps_per_s': steps / el, 'episodes': n}), flush=True)
        env.close()


if __name__ == '__main__':
    which = sys.argv[1:] or ['config2', 'rollout', 'policy', 'fp64', 'zerocopy', 'single']
    for w in which:
        globals()[w]()
