"""CPU-side checks: the C-ABI library loads and exports every symbol include/copter_b200.h
declares (no compute without a GPU), the host shell's non-GPU logic, and the N > 1 sharding
logic over a world_size-2 gloo group."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from gym_copter_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, 'include', 'copter_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b(copter_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 16
    for n in sorted(names):
        assert hasattr(lib, n), n
    m = re.search(r'#define\s+COPTER_ABI_VERSION\s+(\d+)', hdr)
    from gym_copter_b200 import _lib as binding
    assert m and lib.copter_abi_version() == int(m.group(1)) == binding.ABI_VERSION


def test_params_struct_matches_header_and_reference_constants(lib):
    from gym_copter_b200 import default_params, CopterParams
    p = default_params()
    # /root/reference gym_copter/dynamics/vehicles/dji_phantom.py:9-26, dynamics/__init__.py:71-76,
    # envs/task.py:25,32-38, envs/lander.py:17-23
    assert (p.B, p.D, p.M, p.L, p.Ix, p.Iy, p.Iz, p.Jr, p.maxrpm) == (5e-3, 2e-6, 1.38, 0.35, 2, 2, 3, 38e-4, 15000)
    assert (p.landing_vel_x, p.landing_vel_y, p.landing_angle, p.G) == (2.0, 1.0, np.pi / 4, 9.80665)
    assert (p.fps, p.initial_random_force, p.out_of_bounds_penalty, p.max_angle_deg, p.bounds,
            p.initial_altitude, p.max_steps) == (100, 30, 100, 45, 10, 10, 1000)
    assert (p.target_radius, p.yaw_penalty_factor, p.xyz_penalty_factor, p.dz_max, p.dz_penalty,
            p.inside_radius_bonus) == (2, 50, 25, 10, 100, 100)
    assert (p.rho, p.lift_coefficient, p.dynamics_model, p.takeoff_target_altitude) == (1.225, 0.4, 0, 5)
    assert C.sizeof(CopterParams) == 28 * 8 + 8
    assert [lib.copter_obs_size(v) for v in range(7)] == [10, 6, 2, 12, 6, 2, 10]
    assert [lib.copter_action_size(v) for v in range(7)] == [4, 2, 1, 4, 2, 1, 4]
    assert lib.copter_obs_size(7) == -2 and lib.copter_action_size(-1) == -2
    with pytest.raises(TypeError):
        default_params(not_a_field=1)


def test_pid_gains_struct_matches_header_and_reference_constants(lib):
    """CopterPidGains: the ctypes mirror has the header's fields in the header's order, and the
    defaults are the reference's controller constants (attic/mars/pidcontrollers/__init__.py:
    AngularVelocityPidController :124-135, PositionHoldPidController :102-107,
    DescentPidController :110-121, AltitudeHoldPidController :91-99, windup default :14)."""
    from gym_copter_b200 import _lib as binding
    hdr = open(os.path.join(ROOT, 'include', 'copter_b200.h')).read()
    body = re.search(r'typedef struct CopterPidGains \{(.*?)\} CopterPidGains;', hdr, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = [n.strip() for decl in re.findall(r'double\s+([^;]+);', body) for n in decl.split(',')]
    assert names == [f[0] for f in binding.CopterPidGains._fields_] and len(names) == 17
    g = binding.default_pid_gains()
    assert (g.rate_kp, g.rate_ki, g.rate_kd, g.rate_windup, g.rate_big) == (1.0, 0.0, 1.0, 6.0, np.radians(40))
    assert (g.pos_kp, g.pos_ki, g.pos_kd, g.pos_windup, g.pos_target) == (0.00001, 0.1, 4.0, 0.2, 0.0)
    assert (g.descent_kp, g.descent_kd) == (1.15, 1.33)
    assert (g.alt_kp, g.alt_ki, g.alt_kd, g.alt_windup, g.alt_target) == (0.2, 3.0, 0.0, 0.2, 5.0)
    assert binding.SOURCE_KINDS == {'const': 0, 'randn': 1, 'uniform': 2, 'pid': 3, 'pid_hover': 4}
    assert all(re.search(r'COPTER_SRC_%s\s*=\s*%d\b' % (k.upper(), v), hdr) for k, v in binding.SOURCE_KINDS.items())
    with pytest.raises(TypeError):
        binding.default_pid_gains(not_a_gain=1)


def test_argument_validation_without_a_gpu(lib):
    from gym_copter_b200 import default_params
    from gym_copter_b200._lib import CopterBuffers
    p, b = default_params(), CopterBuffers()
    assert lib.copter_step_f32(C.byref(p), C.byref(b), 16, 0, 0, 1, 0, 1, None) == -1      # null buffers
    assert lib.copter_reset_f64(C.byref(p), C.byref(b), 16, 0, 0, None) == -1
    assert lib.copter_step_f32(None, C.byref(b), 16, 0, 0, 1, 0, 1, None) == -1
    bad = default_params(max_steps=4000)
    assert lib.copter_reset_f32(C.byref(bad), C.byref(b), 16, 0, 0, None) == -4


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_fails_loudly_without_gpu():
    import gym_copter_b200 as g
    with pytest.raises(g.CopterError):
        g.LanderVec(8)
    with pytest.raises(g.CopterError):
        g.Dynamics()
    with pytest.raises(g.CopterError):
        g.make('gym_copter:Lander-v0')


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'gym_copter_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert 'copter_oracle' not in src, f


def test_spaces_and_ids():
    from gym_copter_b200.envs import Box, make
    b = Box(-1, 1, (4,))
    assert b.contains(b.sample()) and not b.contains(np.full(4, 2.0, np.float32))
    with pytest.raises(ValueError):
        make('NoSuchEnv-v0')


def test_shard_range_partitions():
    from gym_copter_b200 import shard_range
    for n in (0, 1, 7, 255, 256, 1000, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            if n >= 256 * world:
                assert all(lo % 256 == 0 for lo, _ in r)
                sizes = [hi - lo for lo, hi in r]
                assert max(sizes) - min(sizes) <= 256
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gym_copter_b200.sharding import shard_range, all_reduce_stats
    from oracle.copter_oracle import EnvBatch
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    n_global, steps = 1537, 60
    lo, hi = shard_range(n_global, rank, world)
    # every rank draws the same global action stream and steps only its shard of the oracle
    rng = np.random.default_rng(0)
    env = EnvBatch('Lander3D', hi - lo, seed=11, env_offset=lo)
    env.reset()
    stats = torch.zeros(4, dtype=torch.float64)
    for t in range(steps):
        a = rng.uniform(-1, 1, (n_global, 4))
        obs, r, done, info = env.step(a[lo:hi])
        stats += torch.tensor([done.sum(), r.sum(), info['steps_taken'].sum(), 1.0 if rank == 0 else 0.0], dtype=torch.float64)
    all_reduce_stats(stats)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, env.dyn.x.copy(), env.episode.copy()))
    if rank == 0:
        q.put((stats.tolist(), gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    """world_size 2 on CPU: shards step independently (zero per-step communication), the
    union equals the unsharded batch bit for bit, and the stats all-reduce sums the ranks."""
    import torch.multiprocessing as mp
    from oracle.copter_oracle import EnvBatch
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_global, steps = 1537, 60
    rng = np.random.default_rng(0)
    whole = EnvBatch('Lander3D', n_global, seed=11)
    whole.reset()
    tot = np.zeros(3)
    for t in range(steps):
        obs, r, done, info = whole.step(rng.uniform(-1, 1, (n_global, 4)))
        tot += [done.sum(), r.sum(), info['steps_taken'].sum()]
    assert stats[0] == tot[0] and stats[2] == tot[2] and abs(stats[1] - tot[1]) <= 1e-6 * abs(tot[1])
    assert stats[3] == steps          # only rank 0 contributed to this slot
    for lo, hi, x, ep in gathered:
        assert np.array_equal(x, whole.dyn.x[lo:hi]) and np.array_equal(ep, whole.episode[lo:hi])
    assert sorted((lo, hi) for lo, hi, _, _ in gathered) == [(0, 768), (768, 1537)]


def test_gymnasium_shell_is_guarded_and_registers():
    """Without gymnasium nothing is registered; with a gymnasium-shaped module present (the
    oracle's stand-in is enough: Env, spaces.Box, registration.register) the reference's id
    shape is registered with the reference's time limit."""
    import sys
    from gym_copter_b200 import gym_compat
    had = sys.modules.get('gymnasium')
    try:
        for k in [k for k in sys.modules if k == 'gymnasium' or k.startswith('gymnasium.')]:
            del sys.modules[k]
        sys.modules['gymnasium'] = None          # forces ImportError
        assert gym_compat.register_envs() == []
        del sys.modules['gymnasium']
        from oracle import refshim
        refshim._install_gymnasium_shim()
        ids = gym_compat.register_envs()
        assert 'gym_copter_b200/Lander-v0' in ids and 'gym_copter_b200/Hover3D-v0' in ids and 'gym_copter_b200/Takeoff-v0' in ids and len(ids) == 8
        reg = sys.modules['gymnasium.envs.registration'].registry
        assert reg['gym_copter_b200/Lander-v0']['max_episode_steps'] == 1000
    finally:
        for k in [k for k in sys.modules if k == 'gymnasium' or k.startswith('gymnasium.')]:
            del sys.modules[k]
        if had is not None:
            sys.modules['gymnasium'] = had


def test_vector_env_contract_over_a_stand_in():
    """gymnasium's VectorEnv contract (>= 1.0) as make_vector_env's adapter serves it: same-step autoreset
    declared in metadata, terminal observation under info['final_obs'] with its mask, terminal info under
    info['final_info'], the step limit reported as a truncation and never also as a termination.  The
    wrapped env is a stand-in with CopterVecEnv's return shapes, so this runs without a GPU; the same
    adapter over the real env is exercised in tests/test_gpu_rollout.py."""
    import types
    from gym_copter_b200 import gym_compat
    from gym_copter_b200._lib import CAUSE_TIMEOUT, CAUSE_OOB

    class FakeVec:
        num_envs = 4
        metadata = {'render_modes': ['human'], 'render_fps': 100}
        single_observation_space = single_action_space = observation_space = action_space = object()
        closed = 0

        def reset(self, seed=None, options=None):
            self.seed = seed
            return torch.zeros(4, 10), {}

        def step(self, a):
            cause = torch.tensor([0, CAUSE_TIMEOUT, CAUSE_OOB, 0], dtype=torch.uint8)
            done = torch.tensor([False, True, True, False])
            return (torch.ones(4, 10), torch.arange(4.0), done, (cause & CAUSE_TIMEOUT) != 0,
                    {'cause': cause, 'final_obs': torch.full((4, 10), 7.0)})

        def render(self):
            return None

        def close(self):
            self.closed += 1

        def stats(self):
            return 'forwarded'

    v = gym_compat.VectorEnvAdapter(FakeVec())
    assert v.num_envs == 4 and v.metadata['autoreset_mode'] in ('SameStep', getattr(v.metadata['autoreset_mode'], 'SAME_STEP', 'SameStep'))
    obs, info = v.reset(seed=3)
    assert info == {} and v.env.seed == 3 and obs.shape == (4, 10)
    obs, r, term, trunc, info = v.step(torch.zeros(4, 4))
    assert term.tolist() == [False, False, True, False] and trunc.tolist() == [False, True, False, False]
    assert not (term & trunc).any()
    assert info['_final_obs'].tolist() == [False, True, True, False] and (info['final_obs'][info['_final_obs']] == 7).all()
    assert info['_final_info'].tolist() == [False, True, True, False] and info['final_info']['cause'][2] == CAUSE_OOB
    assert (obs == 1).all()                                   # the NEW episode's first observation (same-step reset)
    assert v.stats() == 'forwarded' and v.unwrapped is v
    v.close(); v.close()
    assert v.env.closed == 1
    # with a gymnasium that has vector.AutoresetMode the enum member is used
    import sys
    fake = types.ModuleType('gymnasium')
    fake.vector = types.SimpleNamespace(AutoresetMode=types.SimpleNamespace(SAME_STEP='enum-member'))
    had = sys.modules.get('gymnasium')
    sys.modules['gymnasium'] = fake
    try:
        assert gym_compat.VectorEnvAdapter(FakeVec()).metadata['autoreset_mode'] == 'enum-member'
    finally:
        if had is not None:
            sys.modules['gymnasium'] = had
        else:
            del sys.modules['gymnasium']


def test_integration_md_stub_matches_the_binding():
    """The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would paste)
    declares the same struct fields, in the same order, as the shipped binding and the header."""
    from gym_copter_b200 import _lib as binding
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    code = re.search(r'```python\n# gym_copter/envs/_b200.py.*?```', doc, flags=re.S).group(0)
    ns = {}
    structs = re.search(r'(class CopterParams\(C\.Structure\):.*?)\nLANDER3D', code, flags=re.S).group(1)
    exec('import ctypes as C\n' + structs, ns)
    for name in ('CopterParams', 'CopterBuffers'):
        mine, theirs = getattr(binding, name), ns[name]
        assert [f[0] for f in theirs._fields_] == [f[0] for f in mine._fields_], name
        assert C.sizeof(theirs) == C.sizeof(mine), name
    # every library call the stub makes exists with that name
    for fn in set(re.findall(r'lib\.(copter_[a-z0-9_]+)\(', code)):
        assert hasattr(binding.load(), fn), fn


def test_profile_facts_come_from_the_committed_captures(tmp_path):
    """bench.py folds numbers of the ncu captures into its line (roofline.traffic, roofline_k4 / k16).  They are
    produced by tools/make_profile_facts.py from the condensed captures: re-deriving them from the committed
    captures must give the committed facts file (so the line and profiles/ describe the same build)."""
    import json
    import subprocess
    import sys
    out = tmp_path / 'facts.json'
    prof = os.path.join(ROOT, 'profiles')
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'make_profile_facts.py'), prof, str(out)], check=True, capture_output=True)
    mine, committed = json.load(open(out)), json.load(open(os.path.join(prof, 'r2_profile_facts.json')))
    for key in ('k4_warp_instr_per_32_env_substeps', 'k16_warp_instr_per_32_env_substeps'):
        assert abs(mine[key] - committed[key]) <= 1e-6 * committed[key], key
    assert mine['k1_traffic']['dram_bytes_per_launch'] == committed['k1_traffic']['dram_bytes_per_launch']
    # sanity of the numbers themselves: DRAM traffic within 5 % of the algorithmic 165 B per env, K = 16 cheaper per substep than K = 4
    assert 0.95 <= mine['k1_traffic']['dram_bytes_per_launch'] / (165.0 * mine['k1_traffic']['envs']) <= 1.05
    assert mine['k16_warp_instr_per_32_env_substeps'] < mine['k4_warp_instr_per_32_env_substeps']
