#!/usr/bin/env python3
"""
bench.py -- env-steps/s of the batched copter step on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched by torchrun)
  python bench.py --impl reference ...                           (CPU arm, rank 0 only)

Workload (BASELINE.json configs[2], per GPU): Lander3D, fp32, 2^24 envs per GPU (weak
scaling; the config's 16M-env batch fits one GPU), same-step auto-reset, action stream
`lander.py --random` (1.625e-2 * N(0,1) per motor, /root/reference lander.py:42) cycled from
a pool of pre-generated action tensors resident in HBM.  A "step" is one launch of the step
kernel over the GPU's whole shard = k_substeps env-steps per env.

One JSON line on stdout (rank 0):
  value     env-steps/s over all ranks, inputs resident in HBM (CUDA events, max over ranks)
  e2e       the same through the host-array API (numpy/pinned host buffers in and out,
            H2D + kernel + D2H inside the timed region)
  roofline  the step kernel against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the per-object Python port (oracle/scalar_port.py, the reference's own
            execution style) on the box's host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'env-steps/sec (whole box) for Lander3D'
UNIT = 'env-steps/s'
STREAMS = ('randn', 'const', 'unif')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs-per-gpu', type=int, default=1 << 24)
    ap.add_argument('--k-substeps', type=int, default=1)
    ap.add_argument('--variant', default='Lander3D')
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
    ap.add_argument('--stream', default='randn', choices=STREAMS)
    ap.add_argument('--pool', type=int, default=4, help='distinct pre-generated action tensors')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--cpu-seconds', type=float, default=6.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the fused-substep side measurements')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every 5 ms)."""

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.stop = index, [], None, set(), False
        self.thread = None

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            try:
                vis = os.environ.get('CUDA_VISIBLE_DEVICES')
                idx = int(vis.split(',')[self.index]) if vis else self.index
            except Exception:
                idx = self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                    'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}

            def loop():
                while not self.stop:
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.reasons.update(k for k, b in bits.items() if r & b)
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.thread is not None:
            self.thread.join(timeout=1)

    def summary(self):
        sm = sorted(self.sm)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.mx,
                'reasons': sorted(self.reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------
def cpu_port_rate(stream, seconds, procs):
    from oracle.scalar_port import run_parallel, run_stream
    if procs <= 1:
        n, el = run_stream(stream, seconds)
        return n / el
    return run_parallel(stream, seconds, procs)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """The reference's CPU implementation of the path: its per-object Python execution style
    (oracle/scalar_port.py; the reference tree itself cannot travel to the GPU box and has no
    native code to compile), one env loop per host core."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_cores()
    # bounded sample: the whole --steps/--warmup run stays near two minutes whatever K and W are
    per_step = max(0.05, min(3.0, 100.0 / max(1, args.steps + args.warmup)))
    from oracle.scalar_port import make_pool, run_parallel, run_stream
    pool = make_pool(cores) if cores > 1 else None
    rates = []
    t0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        if pool is not None:
            r = run_parallel(args.stream, per_step, cores, pool)
        else:
            n, el = run_stream(args.stream, per_step)
            r = n / el
        if i >= args.warmup:
            rates.append(r)
    if pool is not None:
        pool.close()
        pool.join()
    value = sum(rates) / len(rates)
    sample = ('%d processes x one ScalarLander env loop each (oracle/scalar_port.py, the reference\'s per-object '
              'execution style), %s action stream, %.2f s of stepping per bench step, reset on done'
              % (cores, args.stream, per_step))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': per_step * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'Lander3D, reset on done: a bounded sample of the b200 arm\'s workload stepped as single-env '
                               'Python objects on the host cores', 'action_stream': args.stream},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': time.perf_counter() - t0}))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def make_actions(torch, stream, n, a, pool, dtype, device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = []
    for _ in range(pool):
        if stream == 'const':
            t = torch.full((n, a), 1.625e-2, dtype=dtype, device=device)
        elif stream == 'randn':
            t = 1.625e-2 * torch.randn((n, a), dtype=dtype, device=device, generator=g)
        else:
            t = 2 * torch.rand((n, a), dtype=dtype, device=device, generator=g) - 1
        out.append(t.contiguous())
    return out


def bytes_per_launch_per_env(w, a, o):
    """state r+w, packed meta r+w, action read, obs write, reward, done."""
    return 2 * 12 * w + 2 * 4 + a * w + o * 4 + w + 1


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference_arm(args)
    # stdout carries the one JSON line and nothing else: from here on file descriptor 1 is stderr
    # (NCCL's version banner and any other library chatter written to fd 1 end up there), and the
    # line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import gym_copter_b200 as g
    if int(os.environ.get('LOCAL_RANK', '0')) == 0 and not os.path.exists(g._lib.LIB_PATH):
        from gym_copter_b200 import build as _b      # in-tree artefact missing (fresh clone): build it
        _b.build()

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the product has no CPU path)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dtype = torch.float32 if args.dtype == 'f32' else torch.float64
    n, k = args.envs_per_gpu, args.k_substeps
    # headline env: exactly the 165 B/env/launch of SURVEY.md 8(d), no statistics side channel
    env = g.CopterVecEnv(args.variant, n, dtype=dtype, seed=2026, env_offset=rank * n,
                         k_substeps=k, auto_reset=True, track_stats=False)
    A, O, w = env.action_size, env.obs_size, (4 if dtype == torch.float32 else 8)
    actions = make_actions(torch, args.stream, n, A, args.pool, dtype, dev, 1234 + rank)
    env.reset()

    def run(e, steps):
        for i in range(steps):
            e.step(actions[i % len(actions)])

    # ---- device-resident measurement ------------------------------------------------------
    run(env, args.warmup)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = env.launches
    barrier()
    with ClockSampler(local) as clk:
        ev0.record()
        run(env, args.steps)
        ev1.record()
        barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = env.launches - launches0
    value = world * n * k * args.steps / (ms * 1e-3)
    # episode bookkeeping of the same workload, from a statistics-enabled twin of the env (also
    # used for the fused-substep side measurements); its all-reduce is the one optional collective
    senv = g.CopterVecEnv(args.variant, n, dtype=dtype, seed=2026, env_offset=rank * n,
                          k_substeps=k, auto_reset=True, track_stats=True)
    senv.reset()
    run(senv, args.warmup)
    senv.clear_stats()
    sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_stat = min(args.steps, 200)
    sv0.record()
    run(senv, n_stat)
    sv1.record()
    barrier()
    ms_with_stats = max_over_ranks(sv0.elapsed_time(sv1)) / n_stat
    stats = senv.stats(reduce_group=True if world > 1 else None)

    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    b_launch = bytes_per_launch_per_env(w, A, O)
    achieved = b_launch * n / (ms * 1e-3 / args.steps) / 1e9
    traffic = None          # DRAM bytes per launch from the committed ncu --set full capture of this kernel
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')) as f:
            tj = json.load(f)
        if tj['envs'] == n and tj['k_substeps'] == k and args.variant == 'Lander3D' and args.dtype == 'f32':
            traffic = tj['dram_bytes_per_launch']
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'algorithmic_bytes_per_launch': b_launch * n, 'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650',
                'kernel': 'copter_step_kernel<%s,%s>' % (args.dtype, args.variant),
                'bytes_per_env_per_launch': b_launch, 'env_steps_per_launch': n * k}

    # ---- this box's own copy bandwidth, measured like MEASURED_PEAKS.json's hbm_gbs --------
    if not args.no_extras:
        try:
            ca = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
            cb = torch.empty_like(ca)
            ca.fill_(1.0)
            for _ in range(3):
                cb.copy_(ca)
            best = 1e9
            for _ in range(10):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); c0.record(); cb.copy_(ca); c1.record(); torch.cuda.synchronize()
                best = min(best, c0.elapsed_time(c1))
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(50, int(ms / 0.7))          # about as long as the timed region above
            c0.record()
            for _ in range(reps):
                cb.copy_(ca)
            c1.record(); torch.cuda.synchronize()
            gb = 4 * ca.numel() / 1e9
            roofline['copy_here'] = {'burst_gbs': gb / best * 1e3, 'sustained_gbs': gb * reps / c0.elapsed_time(c1) * 1e3,
                                     'how': 'torch b.copy_(a), 1 Gi bf16, read+write bytes; best of 10 / %d back to back' % reps}
            roofline['frac_of_copy_here_sustained'] = achieved / roofline['copy_here']['sustained_gbs']
            del ca, cb
        except Exception as e:
            roofline['copy_here'] = {'unavailable': repr(e)[:200]}

    # ---- fused-substep side measurements (same shard, same stream) -----------------------
    extras = {}
    if not args.no_extras and k == 1:
        for kk in (4, 16):
            senv.k_substeps = kk
            run(senv, 3)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps_kk = max(10, args.steps // 4)
            before = senv.stats()['env_steps']
            e0.record()
            run(senv, steps_kk)
            e1.record()
            barrier()
            ms_kk = max_over_ranks(e0.elapsed_time(e1))
            done_steps = senv.stats()['env_steps'] - before         # idle substeps are not counted
            extras['k%d' % kk] = {'value': world * done_steps / (ms_kk * 1e-3), 'unit': UNIT,
                                  'ms_per_launch': ms_kk / steps_kk,
                                  'note': 'executed env-steps only (envs idle after done within a launch)'}
            # the same K through the host-array API: one action row in, one observation row out per K env-steps
            # (a side number: a failure here, e.g. no page-locked memory left, must not cost the bench line;
            # the collectives stay matched because every rank takes the same path or raises before them)
            try:
                hb = senv.host_buffers()
                hb['action'][:] = actions[0].cpu().numpy()
                senv.step_host(None)
                ok = 1.0
            except Exception:
                ok = 0.0
            if max_over_ranks(1.0 - ok) == 0.0:
                barrier()
                before = senv.stats()['env_steps']
                t0 = time.perf_counter()
                for _ in range(max(3, args.e2e_steps // 4)):
                    senv.step_host(None)
                torch.cuda.synchronize()
                el = max_over_ranks(time.perf_counter() - t0)
                extras['k%d' % kk]['e2e'] = {'value': world * (senv.stats()['env_steps'] - before) / el, 'unit': UNIT,
                                             'api': 'CopterVecEnv.step_host, k_substeps=%d' % kk}
        senv.k_substeps = 1
    del senv
    torch.cuda.empty_cache()

    # ---- configs[4] side measurement: policy-in-the-loop rollout, ONE launch per horizon ---
    policy_rollout = None
    if not args.no_extras and k == 1 and args.variant == 'Lander3D' and args.dtype == 'f32':
        try:
            pn, pT = min(n, 1 << 23), 16
            penv = g.CopterVecEnv('Lander3D', pn, seed=2026, env_offset=rank * pn, write_obs=False)
            penv.reset()
            pol = g.mlp_policy(10, 4, dtype=torch.float32, seed=5)
            pro = g.FusedPolicyRollout(penv, pol.net, pT, out_scale=0.2 * 0.0166, out_offset=0.0166)
            for _ in range(3):
                pro.run()
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            p0.record()
            for _ in range(reps):
                pro.run()
            p1.record()
            barrier()
            pms = max_over_ranks(p0.elapsed_time(p1))
            policy_rollout = {'value': world * pn * pT * reps / (pms * 1e-3), 'unit': UNIT, 'ms_per_env_step': pms / (reps * pT),
                              'workload': 'Lander3D f32, %d envs/GPU, tanh MLP 10-64-64-4 policy + env step fused in '
                                          'copter_policy_rollout_kernel, horizon %d per launch, reward/done rows written' % (pn, pT),
                              'bound': 'MUFU (XU) pipe: 132 tanh per env-step'}
            del pro, penv, pol
        except Exception as e:
            policy_rollout = {'unavailable': repr(e)[:200]}
        torch.cuda.empty_cache()

    # ---- end to end through the host-array API -------------------------------------------
    h = env.host_buffers()
    h['action'][:] = actions[0].cpu().numpy()
    for _ in range(2):
        env.step_host(None)
    barrier()
    t0 = time.perf_counter()
    ret_sum = 0.0
    for i in range(args.e2e_steps):
        obs, rew, dn, _, _ = env.step_host(None)
        ret_sum += float(rew[0])                                    # host read of the result
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {'value': world * n * k * args.e2e_steps / e2e_s, 'unit': UNIT,
           'h2d_bytes_per_step': n * A * w, 'd2h_bytes_per_step': n * (O * 4 + w + 1),
           'ms_per_step': e2e_s / args.e2e_steps * 1e3,
           'api': 'CopterVecEnv.step_host (copter_step_host_%s: chunked H2D + kernel + D2H over 4 streams)' % args.dtype}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        rate = cpu_port_rate(args.stream, args.cpu_seconds, cores)
        one = cpu_port_rate(args.stream, min(3.0, args.cpu_seconds), 1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'single_core': one,
               'sample': '%d processes x one per-object Python Lander loop (oracle/scalar_port.py), %s stream, '
                         '%.0f s, reset on done' % (cores, args.stream, args.cpu_seconds)}
        try:        # context: the oracle's compiled C restatement, OpenMP over all host threads
            from oracle.c_oracle import throughput
            cpu['c_port'] = {'value': throughput(args.stream, 4.0, cores), 'unit': UNIT, 'cores': cores,
                             'single_core': throughput(args.stream, 2.0, 1),
                             'sample': 'oracle/copter_oracle.c, 65536 envs, fp64, %s stream, 4 s' % args.stream}
        except Exception as e:
            cpu['c_port'] = {'unavailable': repr(e)[:200]}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': '%s %s, %d envs/GPU, k_substeps=%d, same-step auto-reset, on-device Philox reset forces'
                                   % (args.variant, args.dtype, n, k),
                       'action_stream': args.stream, 'global_envs': world * n,
                       'l2': 'working set %.2f GB per launch > 126 MB L2, no flush needed' % (b_launch * n / 1e9),
                       'parallelism': 'env shards, one per GPU, no per-step communication'},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
            'clocks': clk.summary(),
            'episodes': dict({kk: stats[kk] for kk in ('episodes', 'mean_length', 'landed', 'crashed', 'oob', 'angle', 'timeout')},
                             ms_per_step_with_statistics=ms_with_stats),
            'fused_substeps': extras,
            'policy_rollout': policy_rollout,
        }
        json_out.write(json.dumps(line) + '\n')
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
