"""
Policy-in-the-loop rollouts (BASELINE.json configs[4]).

The per-step loop the reference's consumers run on the host -- obs -> network -> clip ->
env.step (/root/reference attic/drl/3dtest.py:36-61) -- stays on the device end to end: the
policy reads the env's observation tensor in place, the step kernel reads the policy's action
tensor in place and writes reward / done straight into row t of the [T, N] rollout buffers
(GAE-ready layout), and the whole T-step horizon is captured once in a CUDA graph and replayed,
so a rollout costs one graph launch instead of T x (policy kernels + 1) launches.
"""

import ctypes as C

import torch

from . import _lib
from ._lib import CopterError, VARIANT_IDS


class PolicyRollout:
    """
    env      a CopterVecEnv (already constructed on the target device)
    policy   callable obs[N,O] f32 -> action[N,A] (torch module or function; runs under no_grad)
    horizon  T env.step() calls per rollout
    store_obs  also keep obs_t (the observation the policy saw) in a [T, N, O] buffer
    planar     the policy takes the env's fp32 state planes [3,N,4] in place (see PlanarLinear)
               instead of the packed observation; combine with CopterVecEnv(write_obs=False)
    """

    def __init__(self, env, policy, horizon, store_obs=False, use_cuda_graph=True, planar=False):
        self.env, self.policy, self.horizon = env, policy, int(horizon)
        self.planar = bool(planar)
        if not planar and not env.write_obs:
            raise CopterError('env was built with write_obs=False: use planar=True')
        if planar and store_obs:
            raise CopterError('store_obs needs the packed observation (planar=False)')
        n, dev = env.num_envs, env.device
        self.rewards = torch.zeros((self.horizon, n), dtype=env.dtype, device=dev)
        self.dones = torch.zeros((self.horizon, n), dtype=torch.uint8, device=dev)
        self.obs = torch.zeros((self.horizon, n, env.obs_size), dtype=torch.float32, device=dev) if store_obs else None
        self.last_obs = env.obs                       # observation after the last step (bootstrap value input)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self.launches_per_rollout = self.horizon      # of OUR kernels; the policy's are torch's

    def _body(self):
        env = self.env
        for t in range(self.horizon):
            if self.obs is not None:
                self.obs[t].copy_(env.obs)
            action = self.policy(env.planar_obs() if self.planar else env.obs)
            if action.dtype != env.dtype:
                action = action.to(env.dtype)
            env.step(action, reward_out=self.rewards[t], done_out=self.dones[t])

    @torch.no_grad()
    def run(self):
        """One horizon. Returns (rewards [T,N], dones [T,N] bool view, last_obs [N,O])."""
        if not self.env._is_reset:
            raise CopterError('reset() the env before rolling out')
        if not self.use_cuda_graph:
            self._body()
        else:
            if self._graph is None:
                # warm up on a side stream (allocator + lazy module init), roll the env back,
                # then capture
                snapshot = self.env.state_dict()
                s = torch.cuda.Stream(device=self.env.device)
                s.wait_stream(torch.cuda.current_stream(self.env.device))
                with torch.cuda.stream(s):
                    self._body()
                torch.cuda.current_stream(self.env.device).wait_stream(s)
                torch.cuda.synchronize(self.env.device)
                self.env.load_state_dict(snapshot)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._body()
            self._graph.replay()
            self.env.launches += self.horizon
        return self.rewards, self.dones.view(torch.bool), self.last_obs


class PlanarLinear(torch.nn.Module):
    """
    First layer of a policy that reads the env's fp32 state planes in place instead of a packed
    [N,O] observation: y = sum_p planes[p][:, :w_p] @ W[:, 4p:4p+w_p].T + b, for the leading
    `obs_size` state components (Lander3D: 10 = planes 0, 1 and half of plane 2).  Same result as
    `linear(obs)` up to summation order; saves the observation write in the step kernel.
    """

    def __init__(self, linear, obs_size):
        super().__init__()
        self.linear, self.obs_size = linear, int(obs_size)

    def forward(self, planes):
        w, out = self.linear.weight, None
        for p in range((self.obs_size + 3) // 4):
            k = min(4, self.obs_size - 4 * p)
            x = planes[p] if k == 4 else planes[p][:, :k]
            y = x.to(w.dtype) @ w[:, 4 * p:4 * p + k].t()
            out = y if out is None else out + y
        return out if self.linear.bias is None else out + self.linear.bias


class FusedMLPPolicy:
    """
    The tanh MLP O -> 64 -> 64 -> A evaluated by ONE hand-written kernel (copter_policy_mlp_f32)
    straight from the env's fp32 state planes -- no observation tensor, no [N,64] intermediates in
    HBM.  Two implementations, 16-bit operands (bf16; fp16 for the hidden layers of the tcgen05 kernel) with fp32
    accumulation in both: tcgen05.mma with the
    accumulators in tensor memory, 128 envs per tile (the default; csrc/copter_policy_tc.cuh), and
    warp-level mma.sync with the activations in registers (csrc/copter_policy.cuh; selected by
    COPTER_B200_POLICY_TC=0 in the environment); FusedPolicyRollout has a fused kernel of each kind.  `net` is a torch.nn.Sequential(Linear, Tanh, Linear, Tanh, Linear,
    Tanh) (e.g. mlp_policy(...).net); its weights are read in place, so optimizer updates to
    fp32 parameters are picked up by the next call.  action = out_offset + out_scale * net(obs).
    Use with PolicyRollout(..., planar=True) and CopterVecEnv(write_obs=False).
    """

    def __init__(self, env, net, out_scale=1.0, out_offset=0.0):
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        if len(lin) != 3 or lin[0].out_features != 64 or lin[1].in_features != 64 or lin[1].out_features != 64 \
                or lin[0].in_features != env.obs_size or lin[2].out_features != env.action_size:
            raise CopterError('FusedMLPPolicy needs Linear(O,64), Linear(64,64), Linear(64,A) with tanh activations')
        if env.dtype != torch.float32:
            raise CopterError('FusedMLPPolicy reads the fp32 state planes: build the env with dtype=torch.float32')
        self.env, self.lib = env, _lib.load()
        self.params = []
        for m in lin:
            for p in (m.weight, m.bias):
                if p is None or p.dtype != torch.float32 or p.device != env.device or not p.is_contiguous():
                    raise CopterError('policy parameters must be contiguous fp32 tensors (with bias) on the env device')
                self.params.append(p)
        self.out_scale, self.out_offset = float(out_scale), float(out_offset)
        self.action = torch.zeros((env.num_envs, env.action_size), dtype=torch.float32, device=env.device)
        self.launches = 0

    def __call__(self, planes=None):
        env = self.env
        with torch.cuda.device(env.device):
            _lib.check(self.lib.copter_policy_mlp_f32(
                env.state_planes.data_ptr(), env.num_envs, env.num_envs, VARIANT_IDS[env.variant], 64,
                *[p.data_ptr() for p in self.params], C.c_float(self.out_scale), C.c_float(self.out_offset),
                self.action.data_ptr(), C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)),
                'copter_policy_mlp')
        self.launches += 1
        return self.action


class FusedPolicyRollout:
    """
    The whole policy-in-the-loop horizon as ONE kernel launch (copter_policy_rollout_f32): every
    warp evaluates the tanh MLP O -> 64 -> 64 -> A for its 32 envs from the state its lanes hold
    in registers and steps them, T times, writing only row t of the [T, N] rollout buffers.
    Step for step identical to PolicyRollout(env, FusedMLPPolicy(env, net, ...), T, planar=True).

    env        a float32 CopterVecEnv with k_substeps == 1
    net        torch.nn.Sequential(Linear(O,64), Tanh, Linear(64,64), Tanh, Linear(64,A), Tanh);
               weights are read in place (fp32), so optimizer updates are seen by the next run()
    horizon    T
    store_obs / store_actions   also record obs_t [T,N,O] (what the policy saw) / action_t [T,N,A]
                                (the command before the env's clip), e.g. for a PPO update
    action_std  None (deterministic policy), a float, or an fp32 CUDA tensor [A] (e.g. exp of a learnable
                log-std, read in place): Gaussian exploration, action = policy(obs) + std * N(0,1), the
                noise from Philox keyed by (seed, global env id, env.rollout_step + t) -- reproducible and
                independent of how the rollout is cut into horizons
    """

    def __init__(self, env, net, horizon, out_scale=1.0, out_offset=0.0, store_obs=False, store_actions=False,
                 action_std=None):
        probe = FusedMLPPolicy(env, net, out_scale, out_offset)       # same parameter checks
        if env.k_substeps != 1:
            raise CopterError('FusedPolicyRollout steps with k_substeps == 1')
        self.env, self.lib, self.horizon = env, probe.lib, int(horizon)
        if self.horizon < 1:
            raise CopterError('horizon must be >= 1')
        self.params = probe.params
        if action_std is not None and not isinstance(action_std, torch.Tensor):
            action_std = torch.full((env.action_size,), float(action_std), dtype=torch.float32, device=env.device)
        if action_std is not None and (action_std.dtype != torch.float32 or action_std.device != env.device
                                       or action_std.numel() != env.action_size or not action_std.is_contiguous()):
            raise CopterError('action_std must be a contiguous fp32 tensor [A] on the env device')
        self.action_std = action_std
        self.policy = _lib.CopterMlpPolicy(*[p.data_ptr() for p in self.params], 64, float(out_scale), float(out_offset),
                                           action_std.data_ptr() if action_std is not None else None)
        n, dev = env.num_envs, env.device
        self.rewards = torch.zeros((self.horizon, n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((self.horizon, n), dtype=torch.uint8, device=dev)
        self.obs = torch.zeros((self.horizon, n, env.obs_size), dtype=torch.float32, device=dev) if store_obs else None
        self.actions = torch.zeros((self.horizon, n, env.action_size), dtype=torch.float32, device=dev) if store_actions else None
        self.launches_per_rollout = 1

    def run(self):
        """One horizon. Returns (rewards [T,N], dones [T,N] bool view, last_obs [N,O] or None)."""
        env = self.env
        if not env._is_reset:
            raise CopterError('reset() the env before rolling out')
        with torch.cuda.device(env.device):
            b = env._buffers(None, env._force)
            if not env.write_obs:
                b.obs = None
            _lib.check(self.lib.copter_policy_rollout_f32(
                C.byref(env.params), C.byref(b), C.byref(self.policy), env.num_envs, env.env_offset,
                env.seed_value & 0xFFFFFFFFFFFFFFFF, env.rollout_step, self.horizon, VARIANT_IDS[env.variant],
                _lib.F_AUTO_RESET if env.auto_reset else 0, self.rewards.data_ptr(), self.dones.data_ptr(),
                self.actions.data_ptr() if self.actions is not None else None,
                self.obs.data_ptr() if self.obs is not None else None,
                C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)), 'copter_policy_rollout')
        env.launches += 1
        env.rollout_step += self.horizon
        return self.rewards, self.dones.view(torch.bool), env.obs if env.write_obs else None


def mlp_policy(obs_size, action_size, hidden=64, dtype=torch.bfloat16, device='cuda', seed=0):
    """The small tanh MLP of SURVEY.md 8d config 5 (O -> 64 -> 64 -> A), random init."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    net = torch.nn.Sequential(
        torch.nn.Linear(obs_size, hidden), torch.nn.Tanh(),
        torch.nn.Linear(hidden, hidden), torch.nn.Tanh(),
        torch.nn.Linear(hidden, action_size), torch.nn.Tanh())
    for p in net.parameters():
        p.data = (torch.randn(p.shape, generator=g) * (0.5 / max(1, p.shape[-1]) ** 0.5))
    net = net.to(device=device, dtype=dtype).eval()

    def policy(obs):
        return net(obs.to(dtype)).float()
    policy.net = net
    return policy
