"""
BASELINE INFRASTRUCTURE ONLY -- times the UNMODIFIED reference (simondlevy/gym-copter, loaded from
/root/reference through oracle/refshim.py) on the host cores: one `Lander` env loop per process,
reset() on done, the action streams of SURVEY.md 8(d).  Used by bench.py's CPU legs when the
reference tree is present (the build container); the GPU box has no /root/reference, there the
per-object port (oracle/scalar_port.py, pinned bit-exact to this reference) is timed instead.
"""
import time

import numpy as np

from . import refshim


def available():
    return refshim.reference_available()


def run_stream(kind, seconds, seed=0):
    """Steps one reference Lander for ~`seconds` on the named stream. Returns (env_steps, elapsed_s)."""
    ref = refshim.load_reference()
    rng = np.random.default_rng(seed)
    np.random.seed(seed)                      # the reference draws its reset force from the global numpy RNG
    env = ref.Lander()
    env.reset()
    n, t0 = 0, time.perf_counter()
    while True:
        for _ in range(256):
            if kind == 'const':
                a = 1.625e-2 * np.ones(4)                       # lander.py:21,42
            elif kind == 'randn':
                a = 1.625e-2 * rng.standard_normal(4)           # lander.py:42 --random
            else:
                a = rng.uniform(-1, 1, 4)
            _, _, done, _, _ = env.step(a)
            n += 1
            if done:
                env.reset()
        el = time.perf_counter() - t0
        if el >= seconds:
            return n, el


def _worker(args):
    return run_stream(*args)


def make_pool(procs):
    import multiprocessing as mp
    return mp.get_context('fork').Pool(procs)


def run_parallel(kind, seconds, procs, pool=None):
    """`procs` independent reference env loops, one process each. Returns aggregate env-steps/s."""
    own = pool is None
    if own:
        pool = make_pool(procs)
    try:
        res = pool.map(_worker, [(kind, seconds, 1000 + i) for i in range(procs)], chunksize=1)
    finally:
        if own:
            pool.close()
            pool.join()
    return sum(n / el for n, el in res)
