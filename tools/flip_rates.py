#!/usr/bin/env python3
"""
Developer tool (CPU only): how often does the fp32 path flip a threshold comparison against the fp64
reference arithmetic?  Runs the fp32 host instantiation of the kernels' arithmetic (oracle/copter_host.cpp
-- the GPU reproduces it bit for bit, tests/test_gpu_host_exact.py) against the fp64 C oracle on each
action stream of SURVEY.md 8(d) and writes profiles/r2_fp32_flip_rates.json.  A "flip" is an env whose
done flag, flight status, step counter, episode index -- or reward by more than 1e-4 (an episode ending
one substep apart inside a K-fused launch; the |dz| > 10 shaping penalty moving to the next step) --
differs from the oracle's; it then leaves the comparison.

    python tools/flip_rates.py [envs] [steps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from test_host_restatement import fp32_flip_study      # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    rows = []
    for variant in ('Lander3D', 'Lander2D'):
        for kind in ('const', 'randn', 'hover', 'unif'):
            for k in (1, 4, 16):
                r = fp32_flip_study(variant, kind, k, n, t)
                rows.append(r)
                print('%-9s %-6s K=%-2d episodes %8d flips %5d (%.1e per episode)  state %.1e (by norm %.1e)  reward %.1e'
                      % (variant, kind, k, r['episodes'], r['flips'], r['flips_per_episode'], r['state'], r['state_by_norm'], r['reward']), flush=True)
    out = {'what': 'fp32 host instantiation of gym_copter_b200/csrc/copter_core.h vs the fp64 C oracle (oracle/copter_oracle.c); '
                   'the GPU kernels reproduce the fp32 instantiation bit for bit',
           'metric': '|a - ref| / max(|ref|, 1) per component; by norm: / max(||ref_i||_inf, 1)', 'envs': n, 'steps': t, 'rows': rows}
    with open(os.path.join(ROOT, 'profiles', 'r2_fp32_flip_rates.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
