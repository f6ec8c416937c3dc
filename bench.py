#!/usr/bin/env python3
"""
bench.py -- env-steps/s of the batched copter step on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched by torchrun)
  python bench.py --global-envs 16777216 --gpus 8 ...            (strong scaling: configs[2] as written)
  python bench.py --impl reference ...                           (CPU arm, rank 0 only)

Workload (BASELINE.json configs[2], per GPU): Lander3D, fp32, 2^24 envs per GPU (weak scaling; the
config's 16M-env batch fits one GPU), same-step auto-reset, action stream `lander.py --random`
(1.625e-2 * N(0,1) per motor, /root/reference lander.py:42) cycled from a pool of pre-generated action
tensors resident in HBM.  A "step" is one launch of the step kernel over the GPU's whole shard =
k_substeps env-steps per env.

Before anything is timed the batch is DESYNCHRONISED: every env is rolled ~1000 steps on the device
(copter_rollout, `--random` stream, whose episode lengths spread over 60..400 steps) and then a further
stretch on the bench stream itself, so that the timed region sees the steady-state mix of episode phases:
episodes end, envs reset and Philox forces are drawn inside it (`episodes` in the line; all three
action streams of SURVEY.md 8(d) are measured this way, `streams`).

One JSON line on stdout (rank 0):
  value     env-steps/s over all ranks, inputs resident in HBM: MEDIAN of `--repeats` regions of exactly
            `--steps` launches each (CUDA events, max over ranks); `repeats_ms` lists the regions
  e2e       the same through the host-array API (numpy/pinned host buffers in and out, H2D + kernel +
            D2H inside the timed region), and as a fraction of the bare-copy ceiling of the same bytes
  roofline  the step kernel against the measured HBM peak (MEASURED_PEAKS.json); roofline_k4 / k16:
            the fused-substep launches against min(HBM, instruction issue)
  cpu_baseline  the per-object Python port (oracle/scalar_port.py, the reference's own execution
            style) on the box's host cores, bounded sample; the unmodified reference beside it
            wherever /root/reference exists
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'env-steps/sec (whole box) for Lander3D'
UNIT = 'env-steps/s'
STREAMS = ('randn', 'const', 'unif')
ROLLOUT_SOURCE = {'randn': 'randn', 'const': 'const', 'unif': 'uniform'}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--repeats', type=int, default=5, help='timed regions of --steps launches each; value = median')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs-per-gpu', type=int, default=1 << 24)
    ap.add_argument('--global-envs', type=int, default=0, help='strong scaling: this many envs split over the ranks')
    ap.add_argument('--graph', type=int, default=-1, help='1: replay the launches from a CUDA graph (default: on below 2^22 envs/GPU)')
    ap.add_argument('--k-substeps', type=int, default=1)
    ap.add_argument('--variant', default='Lander3D')
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
    ap.add_argument('--stream', default='randn', choices=STREAMS)
    ap.add_argument('--pool', type=int, default=4, help='distinct pre-generated action tensors')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--cpu-seconds', type=float, default=6.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the side measurements (other streams, fused substeps, policy rollout)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every 5 ms)."""

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.stop = index, [], None, set(), False
        self.power = []
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
                        self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
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
                'reasons': sorted(self.reasons), 'samples': len(sm),
                'power_w_max': max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_port_rate(stream, seconds, procs):
    from oracle.scalar_port import run_parallel, run_stream
    if procs <= 1:
        n, el = run_stream(stream, seconds)
        return n / el
    return run_parallel(stream, seconds, procs)


def cpu_reference_rate(stream, seconds, procs):
    """The unmodified reference, where its tree is present (never on the GPU box)."""
    from oracle import ref_runner
    if not ref_runner.available():
        return None
    if procs <= 1:
        n, el = ref_runner.run_stream(stream, seconds)
        return n / el
    return ref_runner.run_parallel(stream, seconds, procs)


def run_reference_arm(args):
    """The reference's CPU implementation of the path on every host core, one env loop per core: the
    UNMODIFIED reference (oracle/ref_runner.py over /root/reference) where that tree exists, else its
    per-object Python port (oracle/scalar_port.py, bit-exact against the reference; the reference tree
    cannot travel to the GPU box and has no native code to compile)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_cores()
    # bounded sample: the whole --steps/--warmup run stays near two minutes whatever K and W are
    per_step = max(0.05, min(3.0, 100.0 / max(1, args.steps + args.warmup)))
    from oracle import ref_runner
    use_ref = ref_runner.available()
    mod = ref_runner if use_ref else __import__('oracle.scalar_port', fromlist=['x'])
    pool = mod.make_pool(cores) if cores > 1 else None
    rates = []
    t0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        if pool is not None:
            r = mod.run_parallel(args.stream, per_step, cores, pool)
        else:
            n, el = mod.run_stream(args.stream, per_step)
            r = n / el
        if i >= args.warmup:
            rates.append(r)
    if pool is not None:
        pool.close()
        pool.join()
    value = sum(rates) / len(rates)
    what = ('the unmodified reference Lander (/root/reference through oracle/refshim.py)' if use_ref else
            'ScalarLander (oracle/scalar_port.py, the reference\'s per-object execution style)')
    sample = ('%d processes x one env loop each of %s, %s action stream, %.2f s of stepping per bench step, reset on done'
              % (cores, what, args.stream, per_step))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': per_step * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'Lander3D, reset on done: a bounded sample of the b200 arm\'s workload stepped as single-env '
                               'Python objects on the host cores', 'action_stream': args.stream},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'reference' if use_ref else 'port', 'sample': sample},
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


def desynchronise(env, stream):
    """Steady-state mix of episode phases on `stream`: ~1000 device-side steps on the `--random` stream
    (episode lengths spread widely, so the phases decorrelate), then a stretch of the bench stream long
    enough for every env to be inside an episode of that stream."""
    env.reset()
    env.rollout(997, source='randn')
    tail = {'randn': 0, 'const': 800, 'unif': 64}[stream]
    if tail:
        env.rollout(tail, source=ROLLOUT_SOURCE[stream])


def median(v):
    s = sorted(v)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


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
    strong = args.global_envs > 0
    n = (args.global_envs // world // 256 * 256) if strong else args.envs_per_gpu
    k = args.k_substeps
    use_graph = (n < (1 << 22)) if args.graph < 0 else bool(args.graph)
    w = 4 if dtype == torch.float32 else 8

    def measure(stream, kk, steps, repeats, warmup, stats_steps, with_clocks=False):
        """One workload on this rank's shard: desynchronise, warm up, time `repeats` regions of exactly
        `steps` launches.  Returns (list of region ms -- max over ranks, launches, episode statistics of
        the same workload from a statistics-enabled twin, clock summary)."""
        # headline env: exactly the 165 B/env/launch of SURVEY.md 8(d), no statistics side channel
        env = g.CopterVecEnv(args.variant, n, dtype=dtype, seed=2026, env_offset=rank * n, k_substeps=kk,
                             auto_reset=True, track_stats=False)
        actions = make_actions(torch, stream, n, env.action_size, args.pool, dtype, dev, 1234 + rank)
        desynchronise(env, stream)

        def run(e, count, i0=0):
            for i in range(count):
                e.step(actions[(i0 + i) % len(actions)])

        graph = None
        if use_graph:
            # launch-bound regime (small shards): replay `pool` launches per graph node chain so the ~50 us
            # launches are not paced by the Python / ctypes dispatch
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run(env, len(actions))
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    run(env, len(actions))
            torch.cuda.current_stream().wait_stream(side)

        def run_timed(count):
            if graph is None:
                run(env, count)
                return count
            reps = max(1, count // len(actions))
            for _ in range(reps):
                graph.replay()
            return reps * len(actions)

        run_timed(max(warmup, 3))
        regions, launched = [], 0
        launches0 = env.launches
        with ClockSampler(local) as clk:
            for _ in range(repeats):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                ev0.record()
                done = run_timed(steps)
                ev1.record()
                barrier()
                regions.append(max_over_ranks(ev0.elapsed_time(ev1)) * steps / done)     # normalised to `steps` launches
                launched += done
        launches = (env.launches - launches0) if graph is None else launched
        del env
        # episode bookkeeping of the same workload, from a statistics-enabled twin (same seeds, same
        # desynchronisation, same actions); its all-reduce is the one optional collective
        ep, ms_stats = None, None
        if stats_steps:
            senv = g.CopterVecEnv(args.variant, n, dtype=dtype, seed=2026, env_offset=rank * n, k_substeps=kk,
                                  auto_reset=True, track_stats=True)
            desynchronise(senv, stream)
            run(senv, max(warmup, 3))
            senv.clear_stats()
            sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sv0.record()
            run(senv, stats_steps)
            sv1.record()
            barrier()
            ms_stats = max_over_ranks(sv0.elapsed_time(sv1)) / stats_steps
            st = senv.stats(reduce_group=True if world > 1 else None)
            ep = {kk_: st[kk_] for kk_ in ('episodes', 'mean_length', 'landed', 'crashed', 'oob', 'angle', 'timeout', 'env_steps')}
            ep['launches'] = stats_steps
            ep['episodes_per_launch'] = st['episodes'] / stats_steps
            ep['ms_per_step_with_statistics'] = ms_stats
            del senv
        del actions
        torch.cuda.empty_cache()
        return regions, launches, ep, clk.summary()

    # ---- headline: device-resident measurement on --stream -----------------------------------
    regions, launches, episodes, clocks = measure(args.stream, k, args.steps, args.repeats, args.warmup,
                                                  min(args.steps, 200), with_clocks=True)
    ms = median(regions)                                  # ms per region of `steps` launches
    value = world * n * k * args.steps / (ms * 1e-3)
    env0 = g.CopterVecEnv(args.variant, 256, dtype=dtype)
    A, O = env0.action_size, env0.obs_size
    del env0

    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    b_launch = bytes_per_launch_per_env(w, A, O)
    achieved = b_launch * n / (ms * 1e-3 / args.steps) / 1e9
    prof = {}               # numbers read off the committed ncu captures of this build (profiles/r2_profile_facts.json)
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_profile_facts.json')) as f:
            prof = json.load(f)
    except Exception:
        pass
    traffic = None          # DRAM bytes per launch from the committed ncu --set full capture of this kernel
    tj = prof.get('k1_traffic', {})
    if tj and tj.get('envs') == n and k == 1 and args.variant == 'Lander3D' and args.dtype == 'f32':
        traffic = tj.get('dram_bytes_per_launch')
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'algorithmic_bytes_per_launch': b_launch * n,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650',
                'kernel': 'copter_step_kernel<%s,%s>' % (args.dtype, args.variant),
                'bytes_per_env_per_launch': b_launch, 'env_steps_per_launch': n * k,
                'l2_note': ('shard working set %.0f MB <= 126 MB L2: DRAM traffic may fall below the algorithmic bytes'
                            % (b_launch * n / 1e6)) if b_launch * n < 126e6 else None}

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
        torch.cuda.empty_cache()

    # ---- the other action streams of SURVEY.md 8(d), same shard, same procedure ------------------
    streams = {args.stream: {'value': value, 'unit': UNIT, 'ms_per_step': ms / args.steps, 'frac_of_hbm_peak': achieved / peak,
                             'episodes': episodes}}
    if not args.no_extras:
        for s in STREAMS:
            if s == args.stream:
                continue
            rg, _, ep, _ = measure(s, k, args.steps, max(3, args.repeats // 2 + 1), args.warmup, min(args.steps, 200))
            m_ = median(rg)
            streams[s] = {'value': world * n * k * args.steps / (m_ * 1e-3), 'unit': UNIT, 'ms_per_step': m_ / args.steps,
                          'frac_of_hbm_peak': b_launch * n / (m_ * 1e-3 / args.steps) / 1e9 / peak, 'episodes': ep}

    # ---- fused-substep side measurements (same shard, same stream): min(HBM, instruction issue) ----
    extras, roof_k = {}, {}
    if not args.no_extras and k == 1:
        for kk in (4, 16):
            steps_kk = max(10, args.steps // 4)
            rg, _, ep, ck = measure(args.stream, kk, steps_kk, 3, 3, steps_kk)
            m_ = median(rg) / steps_kk                                  # ms per launch
            nominal = world * n * kk / (m_ * 1e-3)
            executed = nominal * ep['env_steps'] / (world * n * kk * ep['launches']) if ep else None
            extras['k%d' % kk] = {'value': executed, 'nominal': nominal, 'unit': UNIT, 'ms_per_launch': m_,
                                  'note': 'value = executed env-steps only (an env that finishes inside a launch idles for the rest of it); '
                                          'nominal counts n*K per launch', 'episodes_per_launch': ep['episodes_per_launch'] if ep else None}
            # min(HBM, pipe) of SURVEY.md 8(d): bytes per env-step fall as 1/K, the issue bound comes from the
            # committed ncu capture of this build (warp-instructions issued per 32 env-substeps) and the SM clock
            # sampled during this very region
            instr = prof.get('k%d_warp_instr_per_32_env_substeps' % kk)
            sm_mhz = ck.get('sm_mhz') or peaks.get('sm_max_mhz') or 1965.0
            hbm_bound = peak * 1e9 * kk / b_launch
            issue_bound = (148 * 4 * sm_mhz * 1e6 / instr * 32) if instr else None
            bound = min(hbm_bound, issue_bound) if issue_bound else hbm_bound
            roof_k['roofline_k%d' % kk] = {
                'bound': 'issue' if issue_bound and issue_bound < hbm_bound else 'hbm', 'achieved': nominal / world, 'unit': 'env-steps/s/GPU (nominal n*K per launch)',
                'hbm_bound': hbm_bound, 'issue_bound': issue_bound, 'peak': bound, 'frac': nominal / world / bound,
                'warp_instr_per_32_env_substeps': instr, 'sm_mhz_during': sm_mhz, 'clock_reasons': ck.get('reasons'),
                'kernel': 'copter_step_kernel<%s,%s>, K-fused loop (straight-line substeps + calm streak)' % (args.dtype, args.variant),
                'how': 'issue bound = 148 SMs x 4 schedulers x SM clock / (warp-instructions executed per 32 env-substeps on the desynchronised '
                       'batch, ncu smsp__inst_executed.sum of this build, profiles/r2_profile_facts.json) x 32'}

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
                                          'copter_policy_rollout_tc_kernel (tcgen05 / TMEM; COPTER_B200_POLICY_ROLLOUT_TC=0 selects the warp-MMA kernel), horizon %d per launch, reward/done rows written' % (pn, pT),
                              'bound': 'MUFU (XU) pipe + issue slots: 104 MUFU tanh and 32 FMA-pipe polynomial tanh per env-step'}
            del pro, penv, pol
        except Exception as e:
            policy_rollout = {'unavailable': repr(e)[:200]}
        torch.cuda.empty_cache()

    # ---- configs[0] side measurement: the reference's own call shape, ONE env, host arrays in and out ---
    single_env = None
    if not args.no_extras and k == 1 and rank == 0:
        try:
            import numpy as np
            senv = g.make('gym_copter:Lander-v0')
            senv.reset()
            sa, s_steps, s_eps = 0.0166 * np.ones(4), 0, 0
            for _ in range(200):
                if senv.step(sa)[2]:
                    senv.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            while s_steps < 3000:
                s_steps += 1
                if senv.step(sa)[2]:
                    senv.reset(); s_eps += 1
            el = time.perf_counter() - t0
            single_env = {'us_per_step': el / s_steps * 1e6, 'steps_per_s': s_steps / el, 'episodes': s_eps,
                          'workload': "make('gym_copter:Lander-v0'): one fp64 env, lander.py's constant command, reset() on done, wall clock "
                                      'around env.step() (numpy in, numpy out: copter_step_host_f64, direct path)'}
            senv.close()
        except Exception as e:
            single_env = {'unavailable': repr(e)[:200]}

    # ---- end to end through the host-array API -------------------------------------------
    env = g.CopterVecEnv(args.variant, n, dtype=dtype, seed=2026, env_offset=rank * n, k_substeps=k, auto_reset=True)
    actions = make_actions(torch, args.stream, n, A, 1, dtype, dev, 1234 + rank)
    desynchronise(env, args.stream)
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
    h2d, d2h = n * A * w, n * (O * 4 + w + 1)
    e2e = {'value': world * n * k * args.e2e_steps / e2e_s, 'unit': UNIT,
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'ms_per_step': e2e_s / args.e2e_steps * 1e3,
           'api': 'CopterVecEnv.step_host (copter_step_host_%s: chunked H2D + kernel + D2H over 4 streams)' % args.dtype}
    # the ceiling of the box's host path for exactly these bytes: plain pinned copies, both directions at once,
    # every rank at the same time (what the PCIe links / the host memory system can move, no kernel at all)
    try:
        ht = env._host_t
        da, dobs = actions[0], env.obs
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def bare(direction):
            barrier()
            t1 = time.perf_counter()
            for _ in range(3):
                if direction in ('h2d', 'both'):
                    with torch.cuda.stream(s_in):
                        da.copy_(ht['action'], non_blocking=True)
                if direction in ('d2h', 'both'):
                    with torch.cuda.stream(s_out):
                        ht['obs'].copy_(dobs, non_blocking=True)
                        ht['reward'].copy_(env.reward, non_blocking=True)
                        ht['done'].copy_(env.done, non_blocking=True)
            torch.cuda.synchronize()
            return max_over_ranks(time.perf_counter() - t1) / 3 * 1e3
        bare('both')
        c_out, c_in, c_both = bare('d2h'), bare('h2d'), bare('both')
        e2e['host_link'] = {'d2h_only_ms': c_out, 'h2d_only_ms': c_in, 'duplex_ms': c_both,
                            'd2h_gbs_per_gpu': d2h / c_out / 1e6, 'h2d_gbs_per_gpu': h2d / c_in / 1e6,
                            'how': 'torch pinned<->device copies of the same tensors on two streams, all %d ranks at once, max over ranks' % world}
        e2e['frac_of_duplex_copy_ceiling'] = c_both / e2e['ms_per_step']
    except Exception as ex:
        e2e['host_link'] = {'unavailable': repr(ex)[:200]}
    # frame-skip through the same API: one command row in, one observation row out per K env-steps
    if not args.no_extras and k == 1:
        for kk in (4, 16):
            try:
                env.k_substeps = kk
                env.step_host(None)
                barrier()
                t1 = time.perf_counter()
                reps = max(3, args.e2e_steps // 2)
                for _ in range(reps):
                    env.step_host(None)
                torch.cuda.synchronize()
                el = max_over_ranks(time.perf_counter() - t1)
                extras.setdefault('k%d' % kk, {})['e2e'] = {'nominal': world * n * kk * reps / el, 'unit': UNIT,
                                                            'api': 'CopterVecEnv.step_host, k_substeps=%d' % kk}
            except Exception as ex:
                extras.setdefault('k%d' % kk, {})['e2e'] = {'unavailable': repr(ex)[:200]}
        env.k_substeps = 1
    del env, actions
    torch.cuda.empty_cache()

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        rate = cpu_port_rate(args.stream, args.cpu_seconds, cores)
        one = cpu_port_rate(args.stream, min(3.0, args.cpu_seconds), 1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'single_core': one,
               'sample': '%d processes x one per-object Python Lander loop (oracle/scalar_port.py), %s stream, '
                         '%.0f s, reset on done' % (cores, args.stream, args.cpu_seconds)}
        try:        # the unmodified reference beside it, wherever its tree exists (never on the GPU box)
            ref_rate = cpu_reference_rate(args.stream, args.cpu_seconds, cores)
            if ref_rate is not None:
                cpu['reference'] = {'value': ref_rate, 'unit': UNIT, 'cores': cores, 'kind': 'reference',
                                    'single_core': cpu_reference_rate(args.stream, min(3.0, args.cpu_seconds), 1),
                                    'sample': '%d processes x one unmodified reference Lander loop (/root/reference), %s stream, %.0f s'
                                              % (cores, args.stream, args.cpu_seconds)}
        except Exception as e:
            cpu['reference'] = {'unavailable': repr(e)[:200]}
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
            'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': '%s %s, %d envs/GPU, k_substeps=%d, same-step auto-reset, on-device Philox reset forces, '
                                   'batch desynchronised before timing (episodes end and reset inside the timed region)'
                                   % (args.variant, args.dtype, n, k),
                       'action_stream': args.stream, 'global_envs': world * n,
                       'l2': ('working set %.2f GB per launch > 126 MB L2, no flush needed' % (b_launch * n / 1e9)) if b_launch * n > 126e6
                             else ('working set %.0f MB per launch fits the 126 MB L2: strong-scaling regime, see roofline.l2_note' % (b_launch * n / 1e6)),
                       'cuda_graph': use_graph,
                       'parallelism': 'env shards, one per GPU, no per-step communication'},
            'repeats_ms': regions, 'timing': 'median of %d regions of exactly %d launches each, CUDA events, max over ranks' % (len(regions), args.steps),
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
            'clocks': clocks,
            'episodes': episodes,
            'streams': streams,
            'fused_substeps': extras,
            'policy_rollout': policy_rollout,
            'single_env': single_env,
        }
        line.update(roof_k)
        json_out.write(json.dumps(line) + '\n')
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
