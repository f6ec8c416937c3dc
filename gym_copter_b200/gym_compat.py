"""
gymnasium-facing shell.  gymnasium is not installed in this image, so
everything here is import-guarded: `register_envs()` is a no-op returning [] when gymnasium is
absent, and with gymnasium present it registers

    gym_copter_b200/Lander-v0        single env, the reference's id shape and time limit
                                     (/root/reference gym_copter/__init__.py:9-13: max_episode_steps=1000)
    gym_copter_b200/<Variant>-v0     the attic-defined variants (Lander2D, Lander1D, Hover3D, ...)

`make_vector_env()` is the batched form under gymnasium's VectorEnv contract (gymnasium >= 1.0):
  * attributes num_envs, single_observation_space, single_action_space, observation_space, action_space;
  * metadata['autoreset_mode'] = SAME_STEP (gymnasium.vector.AutoresetMode when importable, else the string
    'SameStep'): a finished env is reset inside the step that finished it and `obs` is the new episode's
    first observation -- which is exactly the kernels' same-step auto-reset;
  * the terminal observation travels in info['final_obs'] with the boolean mask info['_final_obs'], the
    terminal info (here: the ending cause bits) in info['final_info'] / info['_final_info'];
  * the env's own step limit (envs/task.py:128; gymnasium's TimeLimit(max_episode_steps=1000) for the
    reference, gym_copter/__init__.py:9-13) is reported as `truncations`, every other ending as
    `terminations`, never both.
The adapter adds no arithmetic: it forwards to envs.CopterVecEnv (tensors stay on the device).
"""

VARIANTS = ('Lander3D', 'Lander2D', 'Lander1D', 'Hover3D', 'Hover2D', 'Hover1D', 'Takeoff')


def _gymnasium():
    try:
        import gymnasium
        return gymnasium
    except Exception:
        return None


def _single_entry(variant):
    def make(**kw):
        gym = _gymnasium()
        from .envs import SingleEnv

        class GymSingle(gym.Env):
            metadata = {'render_modes': ['human', 'rgb_array'], 'render_fps': 100}

            def __init__(self, **kw):
                self._env = SingleEnv(variant, **kw)
                self.observation_space = gym.spaces.Box(-float('inf'), float('inf'),
                                                        shape=self._env.observation_space.shape, dtype='float32')
                self.action_space = gym.spaces.Box(-1, 1, self._env.action_space.shape, dtype='float32')
                self.STATE_NAMES, self.TARGET_RADIUS = self._env.STATE_NAMES, self._env.TARGET_RADIUS
                self.FRAMES_PER_SECOND = self._env.FRAMES_PER_SECOND

            def reset(self, seed=None, options=None):
                return self._env.reset(seed=seed, options=options)

            def step(self, action):
                return self._env.step(action)

            def render(self):
                return self._env.render()

            def close(self):
                self._env.close()

            def __getattr__(self, name):            # pose, done, steps, spinning, dynamics, viewer ...
                return getattr(self.__dict__['_env'], name)
        return GymSingle(**kw)
    return make


def register_envs():
    """Registers the ids with gymnasium when it is importable. Returns the ids registered."""
    gym = _gymnasium()
    if gym is None:
        return []
    from gymnasium.envs.registration import register
    ids = []
    for variant in VARIANTS:
        for name in ((variant, 'Lander') if variant == 'Lander3D' else (variant,)):
            env_id = 'gym_copter_b200/%s-v0' % name
            try:
                register(id=env_id, entry_point=_single_entry(variant), max_episode_steps=1000)
                ids.append(env_id)
            except Exception:       # already registered
                pass
    return ids


def _autoreset_same_step():
    gym = _gymnasium()
    mode = getattr(getattr(gym, 'vector', None), 'AutoresetMode', None) if gym is not None else None
    return getattr(mode, 'SAME_STEP', 'SameStep')


class VectorEnvAdapter:
    """gymnasium VectorEnv contract over a CopterVecEnv-shaped object (see the module docstring).
    `make_vector_env` derives it from gymnasium.vector.VectorEnv when that class is importable."""

    def __init__(self, env):
        self.env = env
        self.num_envs = env.num_envs
        for k in ('single_observation_space', 'single_action_space', 'observation_space', 'action_space'):
            setattr(self, k, getattr(env, k))
        self.metadata = dict(getattr(env, 'metadata', {}), autoreset_mode=_autoreset_same_step())
        self.render_mode = None
        self.closed = False

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return obs, dict(info)

    def step(self, actions):
        obs, reward, done, truncated, info = self.env.step(actions)
        out = {}
        cause = info.get('cause')
        if cause is not None:                   # the step limit is TimeLimit's truncation, not a termination
            terminated = done & ~truncated
            out['final_info'] = {'cause': cause}
            out['_final_info'] = done
        else:
            terminated = done
        if 'final_obs' in info:
            out['final_obs'] = info['final_obs']
            out['_final_obs'] = done
        return obs, reward, terminated, truncated, out

    def render(self):
        return self.env.render()

    def close(self, **kw):
        if not self.closed:
            self.env.close()
            self.closed = True

    @property
    def unwrapped(self):
        return self

    def __getattr__(self, name):                # stats(), rollout(), state, ... of the wrapped env
        return getattr(self.__dict__['env'], name)


def make_vector_env(variant='Lander3D', num_envs=1, **kw):
    """The batched env under gymnasium's VectorEnv contract: same-step auto-reset with `final_obs` /
    `final_info`, step-limit endings as truncations.  Subclasses gymnasium.vector.VectorEnv when
    gymnasium is present (so that isinstance checks in trainers pass)."""
    from .envs import CopterVecEnv
    kw.setdefault('auto_reset', True)
    kw.setdefault('keep_final_obs', True)
    kw.setdefault('report_cause', True)
    inner = CopterVecEnv(variant, num_envs, **kw)
    gym = _gymnasium()
    vector = getattr(gym, 'vector', None) if gym is not None else None
    if vector is None or not hasattr(vector, 'VectorEnv'):
        return VectorEnvAdapter(inner)

    class GymCopterVectorEnv(VectorEnvAdapter, vector.VectorEnv):
        pass
    return GymCopterVectorEnv(inner)
