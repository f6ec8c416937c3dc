#!/usr/bin/env python3
"""
Developer tool: A/B builds of the policy kernels (occupancy target, tanh form).  `build`
cross-compiles the variants here into tools/variants/; `run` times each on the GPU box (one
subprocess per variant, selected through COPTER_B200_LIB): the policy kernel alone and the fused
policy + step rollout, Lander3D, 2^23 envs.

    python tools/sweep_policy.py build
    gpurun -- python tools/sweep_policy.py run
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, 'tools', 'variants')
SRC = os.path.join(ROOT, 'gym_copter_b200', 'csrc', 'copter_kernels.cu')

def _v(mt, ctas, *extra):
    return ['-DCOPTER_POLICY_MT=%d' % mt, '-DCOPTER_POLICY_ROLLOUT_CTAS_PER_SM=%d' % ctas, '-DCOPTER_POLICY_CTAS_PER_SM=%d' % ctas] + list(extra)


VARIANTS = {
    'pol_mt2_c4': _v(2, 4), 'pol_mt2_c5': _v(2, 5), 'pol_mt2_c6': _v(2, 6),
    'pol_mt1_c5': _v(1, 5), 'pol_mt1_c6': _v(1, 6), 'pol_mt1_c7': _v(1, 7), 'pol_mt1_c8': _v(1, 8),
    'pol_mt2_c5_bf16x2': _v(2, 5, '-DCOPTER_POLICY_TANH_BF16X2=1'),
}
# FMA-pipe polynomial tanh for the hidden n-tiles in the mask (bit nt), packed f32x2 or scalar
for _m in (0x00, 0x80, 0x88, 0xa8, 0xaa, 0xff):
    VARIANTS['poly_%02x' % _m] = _v(2, 5, '-DCOPTER_POLICY_POLY_MASK=0x%02x' % _m)
VARIANTS['poly_88_scalar'] = _v(2, 5, '-DCOPTER_POLICY_POLY_MASK=0x88', '-DCOPTER_POLICY_POLY_F32X2=0')
VARIANTS['poly_aa_scalar'] = _v(2, 5, '-DCOPTER_POLICY_POLY_MASK=0xaa', '-DCOPTER_POLICY_POLY_F32X2=0')
VARIANTS['poly_88_c4'] = _v(2, 4, '-DCOPTER_POLICY_POLY_MASK=0x88')
VARIANTS['poly_aa_c4'] = _v(2, 4, '-DCOPTER_POLICY_POLY_MASK=0xaa')
if os.environ.get('COPTER_SWEEP_ONLY'):
    VARIANTS = {k: v for k, v in VARIANTS.items() if k.startswith(os.environ['COPTER_SWEEP_ONLY'])}
VARIANTS.update(json.loads(os.environ.get('COPTER_SWEEP_EXTRA', '{}')))


def build():
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        out = os.path.join(VDIR, 'lib_%s.so' % name)
        cmd = ['nvcc', '-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
               '-Xcompiler', '-fPIC', '-shared'] + flags + ['-o', out, SRC]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for n, p in procs:
        o, _ = p.communicate()
        print(n, 'rc', p.returncode, o[-300:] if p.returncode else '')


def time_one():
    import torch
    sys.path.insert(0, ROOT)
    import gym_copter_b200 as g
    n, T = 1 << 23, 16
    env = g.LanderVec(n, seed=3, write_obs=False)
    env.reset()
    pol = g.mlp_policy(10, 4, dtype=torch.float32, seed=5)

    def timed(fn, reps):
        fn(); fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    fused = g.FusedMLPPolicy(env, pol.net, out_scale=0.2 * 0.0166, out_offset=0.0166)
    ms_pol = timed(fused, 30)
    # accuracy against the PyTorch fp32 evaluation, activations pushed into the curved part of tanh
    small = g.LanderVec(1 << 16, seed=3)
    small.reset()
    gen = torch.Generator(device='cuda').manual_seed(0)
    for _ in range(30):
        small.step(0.0166 * (1 + 0.3 * torch.randn((small.num_envs, 4), device='cuda', generator=gen)))
    pol3 = g.mlp_policy(10, 4, dtype=torch.float32, seed=5)
    for p in pol3.net.parameters():
        p.data.mul_(3.0)
    with torch.no_grad():
        err = (g.FusedMLPPolicy(small, pol3.net)() - pol3.net(small.obs)).abs()
    ro = g.FusedPolicyRollout(env, pol.net, T, out_scale=0.2 * 0.0166, out_offset=0.0166)
    ms = timed(ro.run, 8) / T
    print(json.dumps({'policy_kernel_ms': ms_pol, 'rollout_ms_per_env_step': ms, 'rollout_steps_per_s': n / ms * 1e3,
                      'err_max': err.max().item(), 'err_mean': err.mean().item()}))


def run():
    for name in VARIANTS:
        lib = os.path.join(VDIR, 'lib_%s.so' % name)
        if not os.path.exists(lib):
            print(name, 'missing'); continue
        env = dict(os.environ, COPTER_B200_LIB=lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), 'one'], env=env, capture_output=True, text=True, timeout=600)
        print(name, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)


if __name__ == '__main__':
    {'build': build, 'run': run, 'one': time_one}[sys.argv[1]]()
