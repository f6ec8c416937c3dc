"""
gymnasium-facing shell.  gymnasium is not installed in this image, so
everything here is import-guarded: `register_envs()` is a no-op returning [] when gymnasium is
absent, and with gymnasium present it registers

    gym_copter_b200/Lander-v0        single env, the reference's id shape and time limit
                                     (/root/reference gym_copter/__init__.py:9-13: max_episode_steps=1000)
    gym_copter_b200/<Variant>-v0     the attic-defined variants (Lander2D, Lander1D, Hover3D, ...)

and exposes `make_vector_env()` for the batched form with gymnasium's VectorEnv attribute
names (num_envs, single_observation_space, single_action_space, observation_space,
action_space).  The wrappers add no arithmetic: they forward to envs.SingleEnv / envs.CopterVecEnv.
"""

VARIANTS = ('Lander3D', 'Lander2D', 'Lander1D', 'Hover3D', 'Hover2D', 'Hover1D')


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

            @property
            def pose(self):
                return self._env.pose
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


def make_vector_env(variant='Lander3D', num_envs=1, **kw):
    """CopterVecEnv, subclassing gymnasium.vector.VectorEnv when gymnasium is present (so that
    isinstance checks in trainers pass); the plain CopterVecEnv otherwise."""
    from .envs import CopterVecEnv
    gym = _gymnasium()
    vector = getattr(gym, 'vector', None) if gym is not None else None
    if vector is None or not hasattr(vector, 'VectorEnv'):
        return CopterVecEnv(variant, num_envs, **kw)

    class GymCopterVecEnv(CopterVecEnv, vector.VectorEnv):
        def __init__(self, *a, **k):
            CopterVecEnv.__init__(self, *a, **k)
    return GymCopterVecEnv(variant, num_envs, **kw)
