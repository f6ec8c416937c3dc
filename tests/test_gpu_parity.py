"""
GPU parity tests: the CUDA path (through the C ABI, via the package's host shell) against
the oracle and the golden vectors recorded from the reference.

Tolerances (BASELINE.json north_star / SURVEY.md 8d), metric err = |a - ref| / max(|ref|, 1)
per component:
  fp64 path: state / obs / reward <= 1e-9, done flags and step counters bit-exact;
  fp32 path: state / obs <= 1e-4 over 1000 steps, reward <= 1e-4 (same metric).  Discrete outputs:
             the fp32 kernels are bit-exact -- zero flips -- against the fp32 CPU instantiation of
             their own arithmetic (tests/test_gpu_host_exact.py).  Against the fp64 oracle HERE an
             fp32 rounding can flip a threshold comparison one step early/late; that rate is
             measured per action stream on the CPU (tests/test_host_restatement.py, tools/flip_rates.py,
             profiles/r2_fp32_flip_rates.json: 2.2e-3 per episode on the constant-thrust stream, whose
             every episode ends by crossing z = 0; <= 2.4e-4 on the others).  Such envs are counted
             (bound: 0.6 % of episodes, i.e. the measured worst case with margin; every test appends
             its count to gpurun_out/r2_tracker_flips.jsonl) and leave the comparison from that step
             on, since their episode timeline differs.
Saturating action stream.  Actions ~ U(-1,1) command up to 60x hover thrust (SURVEY.md
section 6 "scale note"): accelerations of 1e4 m/s^2, velocity swings of > 100 m/s per step,
episodes of 5-40 steps.  There every fp32 quantity carries an absolute error of 2^-24 times
those magnitudes, so on the fp32 path the state / obs error of such envs is measured against
the size of the env's state vector, |a - ref| / max(||ref_i||_inf, 1) <= 1e-4, and the
per-component figure is only bounded at 1e-3.  The fp64 path keeps the per-component 1e-9.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.copter_oracle import (DynamicsBatch, EnvBatch, VARIANTS, reset_force,   # noqa: E402
                                  STATUS_AIRBORNE, STATUS_CRASHED, STATUS_LANDED, STATUS_LEVELING)

HOVER = 0.016560178212092172
TOL = {torch.float64: 1e-9, torch.float32: 1e-4}
NP_T = {torch.float64: np.float64, torch.float32: np.float32}


def merr(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), 1.0)))


@pytest.fixture(scope='module')
def pkg():
    import gym_copter_b200
    gym_copter_b200.load_library()
    return gym_copter_b200


def action_streams(rng, n, t, a):
    act = np.empty((t, n, a), np.float32)
    for i in range(n):
        kind = i % 4
        if kind == 0:
            act[:, i] = 1.625e-2
        elif kind == 1:
            act[:, i] = 1.625e-2 * rng.standard_normal((t, a))
        elif kind == 2:
            act[:, i] = HOVER * (1 + 0.1 * rng.uniform(-1, 1, (t, a)))
        else:
            act[:, i] = rng.uniform(-1, 1, (t, a))
    return act


class Tracker:
    """Differential comparison with the divergence rule described in the module docstring.
    The discrete outputs (done flag, flight status, step counter, episode index) must agree
    bit for bit; on the fp32 path an env whose discrete outputs differ (a threshold comparison
    flipped by rounding) is counted as a flip and leaves the comparison, because from then on
    its episode timeline is shifted against the oracle's."""

    F32_EPS = 1.2e-7          # float32 observations can differ in the last bit on any path

    def __init__(self, n, dtype, saturating=None):
        self.sync = np.ones(n, bool)
        self.tol, self.exact = TOL[dtype], dtype == torch.float64
        self.flips, self.episodes = 0, 0
        self.max_state = self.max_reward = self.max_obs = self.max_component = 0.0
        # envs driven by the saturating U(-1,1) stream (module docstring); fp32 only
        self.sat = np.zeros(n, bool) if (saturating is None or self.exact) else np.asarray(saturating, bool)

    def _err(self, a, ref, scale):
        a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
        per_comp = np.abs(a - ref) / np.maximum(np.abs(ref), 1.0)
        by_norm = np.abs(a - ref) / np.maximum(scale, 1.0)[:, None]
        self.max_component = max(self.max_component, float(per_comp[self.sync].max(initial=0.0)))
        e = np.where(self.sat[:, None], by_norm, per_comp)
        return float(e[self.sync].max(initial=0.0))

    def compare(self, done, reward, state, obs, discrete, o_done, o_reward, o_state, o_obs, o_discrete):
        bad = done != o_done
        for a, b in zip(discrete, o_discrete):
            bad |= np.asarray(a) != np.asarray(b)
        r_err = np.abs(reward - o_reward) / np.maximum(np.abs(o_reward), 1.0)
        if not self.exact:
            # inside a K-fused launch an episode may end one substep early/late with both sides
            # reporting done: the summed reward then differs by one step's reward
            bad |= o_done & (r_err > self.tol)
        bad &= self.sync
        self.flips += int(bad.sum())
        self.sync &= ~bad
        s = self.sync
        self.episodes += int((o_done & s).sum())
        if s.any():
            scale = np.max(np.abs(np.asarray(o_state, np.float64)), axis=1)
            self.max_reward = max(self.max_reward, float(r_err[s].max()))
            self.max_state = max(self.max_state, self._err(state, o_state, scale))
            self.max_obs = max(self.max_obs, self._err(obs, o_obs, scale))

    def finish(self, min_episodes=1, what=''):
        self.record(what)
        assert self.max_state <= self.tol, self.max_state
        assert self.max_obs <= max(self.tol, self.F32_EPS), self.max_obs
        assert self.max_reward <= self.tol, self.max_reward
        assert self.max_component <= max(10 * self.tol, self.F32_EPS), self.max_component
        assert self.episodes >= min_episodes
        if self.exact:
            assert self.flips == 0
        else:
            assert self.flips <= max(2, 0.006 * self.episodes), (self.flips, self.episodes)

    def record(self, what):
        """One line per comparison into gpurun_out/ (when that directory exists: runs under gpurun)."""
        import inspect
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
        if not os.path.isdir(out):
            return
        caller = what or next((f.function for f in inspect.stack()[2:] if f.function.startswith('test_')), '?')
        with open(os.path.join(out, 'r2_tracker_flips.jsonl'), 'a') as f:
            f.write(json.dumps({'test': caller, 'fp64': self.exact, 'episodes': self.episodes, 'flips': self.flips,
                                'flips_per_episode': self.flips / max(self.episodes, 1), 'max_state': self.max_state,
                                'max_state_per_component': self.max_component, 'max_obs': self.max_obs,
                                'max_reward': self.max_reward}) + '\n')


# ---------------------------------------------------------------------------------------
# golden trajectories recorded from the reference (12 envs x 1000 steps, auto-reset)
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('variant', list(VARIANTS))
def test_golden_trajectories(pkg, golden_dir, variant, dtype):
    g = np.load(os.path.join(golden_dir, 'traj_%s.npz' % variant))
    act = g['action']
    T, N, A = act.shape
    env = pkg.CopterVecEnv(variant, N, dtype=dtype, seed=int(g['seed']), auto_reset=True)
    obs, _ = env.reset()
    tr = Tracker(N, dtype, saturating=np.arange(N) % 4 == 3)
    obs_idx = list(VARIANTS[variant][1])
    # replay the reference's state timeline only at the recorded steps; in between compare
    # flags / rewards / step counters every step
    for t in range(T):
        obs, r, term, trunc, info = env.step(torch.as_tensor(act[t]))
        done = term.cpu().numpy()
        steps = env.steps.cpu().numpy()
        g_steps = np.where(g['done'][t], 1, g['steps'][t])      # after auto-reset the counter is 1
        if t % 10 == 9:
            st = env.state.cpu().numpy()
            ref_st = g['state_every10'][t // 10]
        else:
            st = ref_st = np.zeros((N, 12))
        o = obs.cpu().numpy() if t % 10 == 9 else np.zeros((N, len(obs_idx)), np.float32)
        ref_o = ref_st[:, obs_idx].astype(np.float32) if t % 10 == 9 else o
        tr.compare(done, r.cpu().numpy(), st, o, [steps], g['done'][t], g['reward'][t], ref_st, ref_o, [g_steps])
        assert not trunc.any()
    tr.finish(min_episodes=10)
    if dtype == torch.float64:
        assert merr(env.state.cpu().numpy(), g['final_state']) <= 1e-9


# ---------------------------------------------------------------------------------------
# larger batches against the oracle, on-device Philox resets
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('variant,k', [('Lander3D', 1), ('Lander3D', 4), ('Lander2D', 1), ('Hover3D', 1), ('Lander1D', 2)])
def test_batch_vs_oracle_philox_autoreset(pkg, variant, k, dtype):
    N, T, seed, off = 2048 + 37, 1000 // k, 99, 123456789012
    rng = np.random.default_rng(11)
    act = action_streams(rng, N, T, VARIANTS[variant][2])
    env = pkg.CopterVecEnv(variant, N, dtype=dtype, seed=seed, env_offset=off, k_substeps=k)
    orc = EnvBatch(variant, N, seed=seed, env_offset=off, auto_reset=True)
    obs, _ = env.reset()
    o_obs = orc.reset()
    assert np.array_equal(obs.cpu().numpy(), o_obs)
    tr = Tracker(N, dtype, saturating=np.arange(N) % 4 == 3)
    for t in range(T):
        obs, r, term, _, _ = env.step(torch.as_tensor(act[t]))
        o_obs, o_r, o_done, info = orc.step(act[t].astype(np.float64), k_substeps=k)
        tr.compare(term.cpu().numpy(), r.cpu().numpy(), env.state.cpu().numpy(), obs.cpu().numpy(),
                   [env.steps.cpu().numpy(), env.status.cpu().numpy(), env.episodes.cpu().numpy()],
                   o_done, o_r, orc.dyn.x, o_obs, [orc.steps, orc.dyn.status, orc.episode])
    tr.finish(min_episodes=N)


def test_fp32_1000_step_episodes_within_budget(pkg):
    """Long-lived episodes (near-hover constant and jittered commands): the fp32 path must hold
    1e-4 against the fp64 oracle over the full 1000 steps (the case that breaks an all-fp32
    thrust computation)."""
    N, T = 4096, 999           # the 1000th step ends every episode (task.py:128)
    rng = np.random.default_rng(5)
    base = (HOVER * (1 + 0.0005 * rng.uniform(-1, 1, (N, 4)))).astype(np.float32)
    env = pkg.CopterVecEnv('Lander3D', N, dtype=torch.float32, seed=3, auto_reset=False)
    orc = EnvBatch('Lander3D', N, seed=3, auto_reset=False)
    env.reset(); orc.reset()
    alive = np.ones(N, bool)
    a = torch.as_tensor(base)
    for t in range(T):
        if t % 2:
            jit = (base * (1 + 0.02 * rng.uniform(-1, 1, (N, 4)))).astype(np.float32)
            a_t, a_np = torch.as_tensor(jit), jit
        else:
            a_t, a_np = a, base
        obs, r, term, _, _ = env.step(a_t)
        o_obs, o_r, o_done, _ = orc.step(a_np.astype(np.float64))
        d = term.cpu().numpy()
        alive &= ~(d | o_done)
    assert alive.sum() > N // 8          # plenty of episodes survive the whole horizon
    assert merr(env.state.cpu().numpy()[alive], orc.dyn.x[alive]) <= 1e-4
    assert np.array_equal(env.steps.cpu().numpy()[alive], orc.steps[alive])


# ---------------------------------------------------------------------------------------
# known-answer vectors through the facades
# ---------------------------------------------------------------------------------------

@pytest.fixture(scope='module')
def kat(golden_dir):
    with open(os.path.join(golden_dir, 'kat.json')) as f:
        return json.load(f)


def test_kat1_dynamics_facade(pkg, kat):
    d = pkg.Dynamics(num=3)
    s = np.zeros(12); s[4] = -10
    d.setState(s)
    for k in range(1, 1001):
        d.setMotors(1.625e-2 * np.ones(4))
        g = kat['kat1'].get(str(k))
        if g:
            st = d.getState()
            assert abs(st['z'][1].item() - g['z']) <= 1e-9 and abs(st['dz'][2].item() - g['dz']) <= 1e-9
            assert d.getStatus()[0].item() == g['status']
            assert abs(d.getTime()[0].item() - g['time']) < 1e-12
    assert d.getStatus().tolist() == [STATUS_CRASHED] * 3


def test_kat2_dynamics_facade(pkg, kat):
    g = kat['kat2']
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 1e-5)):
        d = pkg.Dynamics(num=2, dtype=dtype)
        d.setState(g['s0'])
        for k in range(1, 201):
            d.setMotors(g['motors'])
            if str(k) in g['after']:
                assert merr(d.state[0].cpu().numpy(), g['after'][str(k)]) <= tol


def test_kat3_single_env_facade(pkg, kat):
    g = kat['kat3']
    env = pkg.make('gym_copter:Lander-v0')
    obs, info = env.reset(force=g['force'])
    assert obs.dtype == np.float32 and obs.shape == (10,) and info == {}
    rewards = []
    for k in range(1, 1001):
        obs, r, done, trunc, info = env.step(1.625e-2 * np.ones(4))
        assert isinstance(r, float) and isinstance(done, bool) and trunc is False
        rewards.append(r)
        if k == 1:
            assert merr(obs, g['obs1']) <= 1e-7
        if done:
            break
    assert k == g['done_step']
    assert merr(rewards[:3], g['rewards_first3']) <= 1e-9 and rewards[-1] == 0.0
    assert abs(sum(rewards) - g['ret']) <= 1e-9 * abs(g['ret'])
    assert env.vec.status[0].item() == STATUS_CRASHED


def test_soft_landing_fsm(pkg, kat):
    g = kat['soft_landing']
    for dtype in (torch.float64, torch.float32):
        env = pkg.CopterVecEnv('Lander3D', 5, dtype=dtype, auto_reset=False)
        env.reset(force=np.zeros((5, 3)))
        env.set_state(np.tile(np.array(g['s0']), (5, 1)))
        for tr in g['trace']:
            obs, r, term, _, _ = env.step(np.full((5, 4), g['action'], NP_T[dtype]))
            assert env.status.tolist() == [tr['status']] * 5
            assert term.tolist() == [tr['done']] * 5
            assert merr(r.cpu().numpy(), [tr['reward']] * 5) <= TOL[dtype]
            assert merr(env.state.cpu().numpy(), [tr['state']] * 5) <= TOL[dtype]


def test_takeoff_direct_dynamics(pkg, kat):
    for mv, trace in kat['takeoff'].items():
        d = pkg.Dynamics(num=4)
        d.setState(np.zeros(12))
        assert d.getStatus().tolist() == [STATUS_LANDED] * 4
        for k in range(1, 101):
            d.setMotors(float(mv) * np.ones(4))
            for g in trace:
                if g['call'] == k:
                    st = d.getState()
                    assert abs(st['z'][0].item() - g['z']) <= 1e-12 and abs(st['dz'][0].item() - g['dz']) <= 1e-12
                    assert d.getStatus()[0].item() == g['status'] and d._ticks[0].item() == g['ticks']


def test_dynamics_facade_vs_oracle_random(pkg):
    """All four statuses, six-component perturbations, random start states."""
    rng = np.random.default_rng(2)
    N = 1000
    s0 = rng.normal(0, 1, (N, 12)) * np.array([3, 1, 3, 1, 2, 1, .3, .2, .3, .2, .5, .2])
    s0[: N // 4, 4] = np.abs(s0[: N // 4, 4]) * 0.01       # a quarter start touching the ground
    d = pkg.Dynamics(num=N)
    o = DynamicsBatch(N)
    d.setState(s0); o.set_state(s0)
    assert np.array_equal(d.getStatus().cpu().numpy(), o.status)
    for t in range(60):
        m = (HOVER * (1 + 0.5 * rng.uniform(-1, 1, (N, 4)))).astype(np.float32).astype(np.float64)
        if t % 7 == 0:
            f = rng.uniform(-5, 5, (N, 6))
            d.perturb(f); o.set_perturb(f)
        d.setMotors(m); o.set_motors(m)
        assert np.array_equal(d.getStatus().cpu().numpy(), o.status)
        assert np.array_equal(d._ticks.cpu().numpy(), o.ticks)
        assert merr(d.state.cpu().numpy(), o.x) <= 1e-11
        assert merr(d._perturb.cpu().numpy(), o.perturb) <= 1e-12
    assert len(set(o.status.tolist())) >= 3


# ---------------------------------------------------------------------------------------
# properties: sharding / batch-size independence, K-fusion, Philox, edge sizes, errors
# ---------------------------------------------------------------------------------------

def test_reset_force_kernel_bit_exact(pkg):
    lib = pkg.load_library()
    p = pkg.default_params()
    n, off, seed = 5000, (1 << 33) + 17, 0xDEADBEEFCAFE
    ep = torch.randint(0, 2 ** 19, (n,), dtype=torch.int32, device='cuda')
    for fn, td, nd in ((lib.copter_reset_force_f64, torch.float64, np.float64),
                       (lib.copter_reset_force_f32, torch.float32, np.float32)):
        out = torch.zeros((n, 3), dtype=td, device='cuda')
        assert fn(C.byref(p), out.data_ptr(), ep.data_ptr(), n, off, seed, None) == 0
        torch.cuda.synchronize()
        ref = reset_force(seed, np.arange(n, dtype=np.uint64) + np.uint64(off), ep.cpu().numpy(), 30.0, nd)
        assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_sharding_and_batch_size_invariance(pkg, dtype):
    """Results are a function of the global env id only: one shard of 1000 == shards of
    (1, 31, 32, 33, 256, 647) envs with matching env_offset, bit for bit."""
    N, T = 1000, 120
    rng = np.random.default_rng(8)
    act = action_streams(rng, N, T, 4)
    whole = pkg.CopterVecEnv('Lander3D', N, dtype=dtype, seed=42, env_offset=10 ** 6, track_returns=True)
    whole.reset()
    sizes = [1, 31, 32, 33, 256, 647]
    shards, lo = [], 0
    for sz in sizes:
        e = pkg.CopterVecEnv('Lander3D', sz, dtype=dtype, seed=42, env_offset=10 ** 6 + lo, track_returns=True)
        e.reset()
        shards.append((lo, sz, e))
        lo += sz
    assert lo == N
    for t in range(T):
        a = torch.as_tensor(act[t], device='cuda')
        obs, r, term, _, _ = whole.step(a)
        for lo, sz, e in shards:
            o2, r2, t2, _, _ = e.step(a[lo:lo + sz])
            assert torch.equal(o2, obs[lo:lo + sz]) and torch.equal(r2, r[lo:lo + sz]) and torch.equal(t2, term[lo:lo + sz])
    for lo, sz, e in shards:
        assert torch.equal(e.state, whole.state[lo:lo + sz]) and torch.equal(e.meta, whole.meta[lo:lo + sz])
    tot = whole.stats()
    parts = [e.stats() for _, _, e in shards]
    for k in ('episodes', 'length_sum', 'landed', 'crashed', 'oob', 'angle', 'timeout', 'env_steps'):
        assert tot[k] == sum(p[k] for p in parts)
    assert abs(tot['return_sum'] - sum(p['return_sum'] for p in parts)) <= 1e-6 * max(1, abs(tot['return_sum']))
    assert tot['episodes'] > 100 and tot['env_steps'] == N * T


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_k_fused_equals_single_steps_when_nothing_finishes(pkg, dtype):
    N, K = 513, 8
    rng = np.random.default_rng(4)
    a = torch.as_tensor((HOVER * (1 + 0.01 * rng.uniform(-1, 1, (N, 4)))).astype(np.float32))
    e1 = pkg.CopterVecEnv('Lander3D', N, dtype=dtype, seed=1, k_substeps=K)
    e2 = pkg.CopterVecEnv('Lander3D', N, dtype=dtype, seed=1, k_substeps=1)
    e1.reset(); e2.reset()
    for it in range(10):
        o1, r1, d1, _, _ = e1.step(a)
        rs = torch.zeros_like(r1)
        for k in range(K):
            o2, r2, d2, _, _ = e2.step(a)
            rs += r2
            assert not d2.any()
        assert not d1.any()
        assert torch.equal(e1.state, e2.state) and torch.equal(o1, o2) and torch.equal(e1.meta, e2.meta)
        assert merr(r1.cpu().numpy(), rs.cpu().numpy()) <= (1e-12 if dtype == torch.float64 else 2e-5)


def test_stats_and_final_obs_vs_oracle(pkg):
    N, T, seed = 777, 400, 5
    rng = np.random.default_rng(9)
    act = action_streams(rng, N, T, 4)
    env = pkg.CopterVecEnv('Lander3D', N, dtype=torch.float64, seed=seed, track_returns=True, keep_final_obs=True)
    orc = EnvBatch('Lander3D', N, seed=seed)
    env.reset(); orc.reset()
    ep = ret_sum = len_sum = 0
    run = np.zeros(N)
    cause_counts = np.zeros(6)
    for t in range(T):
        pre = orc.dyn.x.copy()
        obs, r, term, _, info = env.step(torch.as_tensor(act[t]))
        o_obs, o_r, o_done, o_info = orc.step(act[t].astype(np.float64))
        run += o_r
        ep += o_done.sum(); ret_sum += run[o_done].sum(); len_sum += (o_info['final_steps'][o_done] - 1).sum()
        run[o_done] = 0
        for b in range(6):
            cause_counts[b] += ((o_info['cause'] >> b) & 1).sum()
        assert np.array_equal(term.cpu().numpy(), o_done)
    s = env.stats()
    assert s['episodes'] == ep and s['length_sum'] == len_sum and s['env_steps'] == N * T
    assert abs(s['return_sum'] - ret_sum) <= 1e-9 * abs(ret_sum)
    assert [s[k] for k in ('landed', 'bonus', 'oob', 'angle', 'crashed', 'timeout')] == list(cause_counts)
    # terminal observation: envs that finish on the last step have final_obs == obs of the
    # terminal state, which for an over-angle/oob/crash event is NOT the reset observation
    fo = info['final_obs'].cpu().numpy()
    d = term.cpu().numpy()
    assert d.any() and not np.array_equal(fo[d], obs.cpu().numpy()[d])


@pytest.mark.parametrize('n', [1, 2, 31, 32, 33, 255, 256, 257, 1025])
def test_edge_sizes(pkg, n):
    """Partial warps / partial tiles, every env on the saturating stream (reset-dominated)."""
    env = pkg.CopterVecEnv('Lander3D', n, dtype=torch.float32, seed=7)
    orc = EnvBatch('Lander3D', n, seed=7)
    env.reset(); orc.reset()
    rng = np.random.default_rng(n)
    tr = Tracker(n, torch.float32, saturating=np.ones(n, bool))
    for t in range(30):
        a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        obs, r, term, _, _ = env.step(a)
        o_obs, o_r, o_done, _ = orc.step(a.astype(np.float64))
        tr.compare(term.cpu().numpy(), r.cpu().numpy(), env.state.cpu().numpy(), obs.cpu().numpy(),
                   [env.steps.cpu().numpy(), env.status.cpu().numpy()], o_done, o_r, orc.dyn.x, o_obs,
                   [orc.steps, orc.dyn.status])
    tr.finish(min_episodes=1 if n < 4 else n)


def test_abi_argument_errors(pkg):
    lib = pkg.load_library()
    from gym_copter_b200._lib import CopterBuffers
    p = pkg.default_params()
    env = pkg.CopterVecEnv('Lander3D', 64)
    env.reset()
    a = torch.zeros((64, 4), device='cuda')
    b = env._buffers(a)
    assert lib.copter_step_f32(C.byref(p), C.byref(b), 0, 0, 0, 1, 0, 1, None) == 0          # empty shard
    assert lib.copter_step_f32(C.byref(p), C.byref(b), 64, 0, 0, 1, 17, 1, None) == -2       # variant
    assert lib.copter_step_f32(C.byref(p), C.byref(b), 64, 0, 0, 0, 0, 1, None) == -4        # k < 1
    assert lib.copter_step_f32(C.byref(p), C.byref(b), -5, 0, 0, 1, 0, 1, None) == -4
    b2 = env._buffers(a)
    b2.state = env.state_planes.data_ptr() + 4
    assert lib.copter_step_f32(C.byref(p), C.byref(b2), 64, 0, 0, 1, 0, 1, None) == -3       # alignment
    b3 = env._buffers(None)
    assert lib.copter_step_f32(C.byref(p), C.byref(b3), 64, 0, 0, 1, 0, 1, None) == -1       # missing action
    bad = pkg.default_params(max_steps=5000)            # past the 11-bit steps field: needs wide counters (meta_hi)
    assert lib.copter_step_f32(C.byref(bad), C.byref(b), 64, 0, 0, 1, 0, 1, None) == -4
    wide = pkg.CopterVecEnv('Lander3D', 64, max_steps=5000)
    wide.reset()
    assert wide.wide and lib.copter_step_f32(C.byref(bad), C.byref(wide._buffers(a)), 64, 0, 0, 1, 0, 1, None) == 0
    with pytest.raises(pkg.CopterError):
        pkg.CopterVecEnv('Lander3D', 8).step(torch.zeros(8, 4))                              # step before reset
    with pytest.raises(ValueError):
        env.step(torch.zeros(63, 4))


def test_non_default_task_params(pkg):
    kw = dict(initial_altitude=4.0, max_steps=50, bounds=3.0, initial_random_force=10.0)
    from oracle.copter_oracle import OracleParams
    N = 300
    env = pkg.CopterVecEnv('Lander3D', N, dtype=torch.float64, seed=2, **kw)
    orc = EnvBatch('Lander3D', N, params=OracleParams(**kw), seed=2)
    env.reset(); orc.reset()
    rng = np.random.default_rng(1)
    n_done = 0
    for t in range(160):
        a = (HOVER * (1 + 0.05 * rng.uniform(-1, 1, (N, 4)))).astype(np.float32)
        obs, r, term, _, _ = env.step(a)
        o_obs, o_r, o_done, _ = orc.step(a.astype(np.float64))
        assert np.array_equal(term.cpu().numpy(), o_done) and merr(r.cpu().numpy(), o_r) <= 1e-9
        n_done += o_done.sum()
    assert merr(env.state.cpu().numpy(), orc.dyn.x) <= 1e-9 and n_done >= 2 * N


def test_checkpoint_roundtrip(pkg):
    N = 500
    rng = np.random.default_rng(6)
    act = action_streams(rng, N, 60, 4)
    e1 = pkg.CopterVecEnv('Lander3D', N, seed=9)
    e1.reset()
    for t in range(30):
        e1.step(act[t])
    ck = e1.state_dict()
    e2 = pkg.CopterVecEnv('Lander3D', N, seed=0)
    e2.load_state_dict(ck)
    for t in range(30, 60):
        o1, r1, d1, _, _ = e1.step(act[t])
        o2, r2, d2, _, _ = e2.step(act[t])
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)


# ---------------------------------------------------------------------------------------
# BASELINE.json sizes: size-independent properties on a 2^22-env shard
# ---------------------------------------------------------------------------------------

def test_full_size_sampled_parity_and_conservation(pkg):
    """4 M envs (config 3's per-GPU shard is 2 M): (a) 4096 randomly chosen envs, re-run alone
    by the C oracle from their global ids and their rows of the action tensors, agree with the
    big batch; (b) episode bookkeeping is conserved: sum of done flags == episodes statistic,
    sum over envs of the episode counters == the same, executed env-steps == N * T;
    (c) stepping the second half of the batch as its own shard gives bit-identical results."""
    from oracle.c_oracle import CEnvBatch
    N, T, seed = 1 << 22, 120, 31337
    g = torch.Generator(device='cuda').manual_seed(1)
    env = pkg.CopterVecEnv('Lander3D', N, dtype=torch.float32, seed=seed, track_stats=True)
    half = pkg.CopterVecEnv('Lander3D', N // 2, dtype=torch.float32, seed=seed, env_offset=N // 2)
    idx = np.sort(np.random.default_rng(0).choice(N, 4096, replace=False))
    idx_t = torch.as_tensor(idx, device='cuda')
    orc = CEnvBatch('Lander3D', len(idx), seed=seed, env_ids=idx)
    env.reset(); half.reset(); orc.reset()
    done_total = 0
    sync = np.ones(len(idx), bool)
    worst = 0.0
    for t in range(T):
        a = 1.625e-2 * torch.randn((N, 4), device='cuda', generator=g)
        if t % 4 == 0:
            a[::5] = 2 * torch.rand((len(a[::5]), 4), device='cuda', generator=g) - 1
        obs, r, term, _, _ = env.step(a)
        o2, r2, t2, _, _ = half.step(a[N // 2:])
        assert torch.equal(o2, obs[N // 2:]) and torch.equal(r2, r[N // 2:]) and torch.equal(t2, term[N // 2:])
        done_total += int(term.sum().item())
        o_obs, o_r, o_done, _ = orc.step(a[idx_t].cpu().numpy().astype(np.float64))
        d = term[idx_t].cpu().numpy()
        sync &= (d == o_done) & (env.status[idx_t].cpu().numpy() == orc.status)
        x = env.state[idx_t].cpu().numpy()
        scale = np.maximum(np.abs(orc.x).max(1, keepdims=True), 1.0)          # saturating commands present
        worst = max(worst, float((np.abs(x - orc.x) / scale)[sync].max()),
                    float((np.abs(r[idx_t].cpu().numpy() - o_r) / np.maximum(np.abs(o_r), 1))[sync].max()))
    assert sync.sum() >= 0.99 * len(idx) and worst <= 1e-4, (sync.sum(), worst)
    s = env.stats()
    assert s['episodes'] == done_total == int(env.episodes.to(torch.int64).sum().item())
    assert s['env_steps'] == N * T
    assert torch.equal(half.state, env.state[N // 2:]) and torch.equal(half.meta, env.meta[N // 2:])


def test_quarter_billion_envs_64bit_indexing(pkg):
    """2^28 + 77 Lander3D envs on one GPU (31 GB of the 180 GB; the observation tensor alone has
    2.7e9 elements, past 2^31): 24 fused rollout steps on the U(-1,1) source (bit-exact commands
    in the oracle, ~7-step episodes, so resets are exercised) and one step() launch, checked for
    env ids sampled at the head, around the 2^31-element boundaries of the obs / state / action
    indexing, and in the ragged tail, against the oracle run alone on those global ids."""
    from oracle.copter_oracle import source_actions
    n, T, seed = (1 << 28) + 77, 24, 4242
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip('needs 40 GB of free device memory')
    edges = [0, (1 << 31) // 12, (1 << 31) // 10, (1 << 31) // 4, 1 << 27, 1 << 28]
    idx = np.unique(np.concatenate([np.arange(max(e - 40, 0), min(e + 40, n)) for e in edges] + [np.arange(n - 100, n)]))
    idx_t = torch.as_tensor(idx, device='cuda')
    env = pkg.CopterVecEnv('Lander3D', n, dtype=torch.float32, seed=seed)
    orc = EnvBatch('Lander3D', len(idx), seed=seed, env_ids=idx)
    env.reset(); orc.reset()
    assert np.array_equal(env.obs[idx_t].cpu().numpy(), orc.observe())
    out = env.rollout(T, source='uniform', record_dones=True)
    dones = out['dones'][:, idx_t].cpu().numpy()
    sync = np.ones(len(idx), bool)
    for t in range(T):
        a = source_actions(seed, idx, t, 'uniform', 1.0, 0.0, 4, np.float32).astype(np.float64)
        _, _, o_done, _ = orc.step(a)
        sync &= dones[t] == o_done
    a = torch.full((n, 4), 0.0166, device='cuda')
    obs, r, term, _, _ = env.step(a)
    o_obs, o_r, o_done, _ = orc.step(np.full((len(idx), 4), np.float64(np.float32(0.0166))))
    mw = env.meta[idx_t].to(torch.int64).cpu().numpy() & 0xFFFFFFFF                # sampled rows only: no full-size temporaries
    sync &= (term[idx_t].cpu().numpy() == o_done) & ((mw & 3) == orc.dyn.status)
    x = env.state_planes[:, idx_t, :].permute(1, 0, 2).reshape(len(idx), 12).cpu().numpy()
    scale = np.maximum(np.abs(orc.dyn.x).max(1, keepdims=True), 1.0)              # saturating commands
    worst = float((np.abs(x - orc.dyn.x) / scale)[sync].max())
    worst_o = float((np.abs(obs[idx_t].cpu().numpy() - o_obs) / scale)[sync].max())
    assert sync.sum() >= 0.97 * len(idx) and worst <= 1e-4 and worst_o <= 1e-4, (sync.sum(), len(idx), worst, worst_o)
    assert (((mw >> 2) & 2047)[sync] == orc.steps[sync]).all()
    del env, a, out
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------
# randomised shapes (hypothesis): any n / k / shard offset / variant / seed
# ---------------------------------------------------------------------------------------

def test_random_shapes_vs_oracle(pkg):
    from hypothesis import given, settings, strategies as st, HealthCheck

    # derandomize: the driver's run draws the same examples every time (COPTER_HYP_EXAMPLES=500 COPTER_HYP_RANDOM=1
    # python -m pytest ... is the exploratory run); a statistical bound that a fresh draw can exceed is no test
    @settings(max_examples=int(os.environ.get('COPTER_HYP_EXAMPLES', '25')), deadline=None, suppress_health_check=list(HealthCheck),
              derandomize=os.environ.get('COPTER_HYP_RANDOM', '0') != '1', database=None)
    @given(n=st.integers(1, 700), k=st.integers(1, 7), off=st.integers(0, 2 ** 40),
           seed=st.integers(0, 2 ** 64 - 1), variant=st.sampled_from(list(VARIANTS)),
           f64=st.booleans(), auto=st.booleans())
    def check(n, k, off, seed, variant, f64, auto):
        dtype = torch.float64 if f64 else torch.float32
        env = pkg.CopterVecEnv(variant, n, dtype=dtype, seed=seed, env_offset=off, k_substeps=k, auto_reset=auto)
        orc = EnvBatch(variant, n, seed=seed, env_offset=off, auto_reset=auto)
        assert np.array_equal(env.reset()[0].cpu().numpy(), orc.reset())
        rng = np.random.default_rng(seed % 2 ** 32)
        sat = rng.random(n) < 0.3                   # these envs get U(-1,1) commands throughout; the others stay near hover
        tr = Tracker(n, dtype, saturating=sat)      # and are held to the per-component 1e-4
        for t in range(12):
            a = np.where(sat[:, None], rng.uniform(-1, 1, (n, env.action_size)),
                         HOVER * (1 + 0.2 * rng.standard_normal((n, env.action_size)))).astype(np.float32)
            obs, r, term, _, _ = env.step(a)
            o_obs, o_r, o_done, _ = orc.step(a.astype(np.float64), k_substeps=k)
            tr.compare(term.cpu().numpy(), r.cpu().numpy(), env.state.cpu().numpy(), obs.cpu().numpy(),
                       [env.steps.cpu().numpy(), env.status.cpu().numpy(), env.episodes.cpu().numpy()],
                       o_done, o_r, orc.dyn.x, o_obs, [orc.steps, orc.dyn.status, orc.episode])
        tr.record('test_random_shapes_vs_oracle')
        assert tr.max_state <= tr.tol and tr.max_reward <= tr.tol and tr.max_obs <= max(tr.tol, tr.F32_EPS), (tr.max_state, tr.max_reward, tr.max_obs)
        # fp64: no flips.  fp32: a rounding may move a threshold comparison by one step -- the same allowance as
        # Tracker.finish (measured rates: profiles/r2_fp32_flip_rates.json; the saturating third of these envs ends an
        # episode every 5-40 steps, so up to ~1500 episodes per example)
        assert tr.flips == 0 if f64 else tr.flips <= max(2, 0.006 * tr.episodes), (tr.flips, tr.episodes, n, k, variant)

    check()


# ---------------------------------------------------------------------------------------
# alternate vehicle / world model (attic/mars): lift-model thrust, gyroscopic Omega, Mars G / rho
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('model', [1, 2, 3])
def test_mars_model_vs_oracle(pkg, model, dtype):
    from oracle.copter_oracle import OracleParams
    world = dict(G=3.721, rho=0.017, lift_coefficient=0.4, dynamics_model=model)
    N, T = 1500, 400
    S = .05 * 0.35 * 4
    if model & 1:
        hover_w = np.sqrt(1.38 * world['G'] / 4 / (0.5 * world['rho'] * S * 0.4)) / (0.35 / 2)
    else:
        hover_w = np.sqrt(1.38 * world['G'] / 4 / 5e-3)
    hover = hover_w / (15000 * np.pi / 30)
    env = pkg.CopterVecEnv('Hover3D', N, dtype=dtype, seed=21, **world)
    orc = EnvBatch('Hover3D', N, params=OracleParams(**world), seed=21)
    env.reset(); orc.reset()
    rng = np.random.default_rng(model)
    tr = Tracker(N, dtype)
    for t in range(T):
        a = hover * (1 + 0.15 * rng.uniform(-1, 1, (N, 4)))
        a[::3, 1:3] *= 1.6                                   # a third of the fleet rolls over
        a = a.astype(np.float32)
        obs, r, term, _, _ = env.step(a)
        o_obs, o_r, o_done, _ = orc.step(a.astype(np.float64))
        tr.compare(term.cpu().numpy(), r.cpu().numpy(), env.state.cpu().numpy(), obs.cpu().numpy(),
                   [env.steps.cpu().numpy(), env.status.cpu().numpy()], o_done, o_r, orc.dyn.x, o_obs,
                   [orc.steps, orc.dyn.status])
    tr.finish(min_episodes=0)
    assert np.abs(orc.dyn.x[:, 7]).max() > 1e-3
    # and through the Dynamics facade (take-off from the ground under Mars gravity)
    d = pkg.Dynamics(params=world, num=64, dtype=dtype)
    o = DynamicsBatch(64, OracleParams(**world), np.float64)
    d.setState(np.zeros(12)); o.set_state(np.zeros((64, 12)))
    for t in range(150):
        m = np.clip(1.3 * hover * (1 + 0.05 * rng.uniform(-1, 1, (64, 4))), 0, 1).astype(np.float32).astype(np.float64)
        d.setMotors(m); o.set_motors(m)
        assert np.array_equal(d.getStatus().cpu().numpy(), o.status)
    assert merr(d.state.cpu().numpy(), o.x) <= TOL[dtype] and (o.status == STATUS_AIRBORNE).all()


def test_integration_md_stub_replays_kat3(pkg, kat):
    """The reference-side ctypes stub printed in INTEGRATION.md, executed as it stands (only the
    library path is filled in) and driven through KAT-3."""
    import re
    from gym_copter_b200 import _lib as binding
    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'INTEGRATION.md')).read()
    code = re.search(r'```python\n(# gym_copter/envs/_b200.py.*?)```', doc, flags=re.S).group(1)
    code = code.replace("C.CDLL('libcopter_b200.so')", 'C.CDLL(%r)' % binding.LIB_PATH)
    ns = {}
    exec(code, ns)
    g = kat['kat3']
    env = ns['B200Lander'](num_envs=3)
    obs, info = env.reset(force=np.tile(np.asarray(g['force'], np.float64)[:3], (3, 1)))
    assert obs.dtype == np.float32 and obs.shape == (3, 10)
    rewards = []
    for k in range(1, 1001):
        obs, r, done, trunc, info = env.step(np.tile(1.625e-2 * np.ones(4), (3, 1)))
        rewards.append(r.copy())
        if k == 1:
            assert merr(obs[1], g['obs1']) <= 1e-7
        if done.any():
            break
    assert k == g['done_step'] and done.all()
    rewards = np.array(rewards)
    assert merr(rewards[:3, 0], g['rewards_first3']) <= 1e-9 and (rewards[-1] == 0.0).all()
    assert abs(rewards[:, 2].sum() - g['ret']) <= 1e-9 * abs(g['ret'])
    # a reset() per episode draws a new force each time (task.py:175-184)
    env2 = ns['B200Lander'](num_envs=4)
    firsts = []
    for _ in range(3):
        env2.reset()
        firsts.append(env2.step(np.tile(0.0166 * np.ones(4), (4, 1)))[0][:, 1].copy())
    assert not np.array_equal(firsts[0], firsts[1]) and not np.array_equal(firsts[1], firsts[2])
