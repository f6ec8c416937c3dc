#!/usr/bin/env python3
"""
Developer tool: reads the condensed `ncu --set full` captures of the step kernel (r2_step_kernel_k{1,4,16}_ncu_full.txt,
written by tools/_ncu_capture.sh) and writes the numbers bench.py folds into its line -- DRAM bytes per launch at K = 1
(`roofline.traffic`) and the executed warp-instructions per 32 env-substeps at K = 4 / 16 (`roofline_k4` / `roofline_k16`)
-- as profiles/r2_profile_facts.json.  Run on the GPU box right after the captures and before bench.py, so that the
line and the captures describe the same build.

    python tools/make_profile_facts.py [capture_dir = gpurun_out] [out = profiles/r2_profile_facts.json]
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3,
        'hz': 1e-9, 'Khz': 1e-6, 'Mhz': 1e-3, 'Ghz': 1.0}


def first_launch(path):
    """{metric: (value, unit)} of launch 0 of a condensed capture"""
    out, seen = {}, False
    for line in open(path):
        if line.startswith('== launch'):
            if seen:
                break
            seen = True
            out['kernel'] = line.split(':', 1)[1].strip()
        elif seen:
            m = re.match(r'(\S+)\s+([-0-9.eE+]+)\s*(\S*)', line)
            if m:
                out[m.group(1)] = (float(m.group(2)), m.group(3))
    return out


def scaled(m, key):
    v, u = m[key]
    return v * UNIT.get(u, 1.0)


def main(src, dst):
    envs = 1 << 24                                  # tools/profile_k.py's batch
    facts = {'note': 'numbers read off the ncu --set full captures of this build by tools/make_profile_facts.py '
                     '(tools/_gpu_job.sh -> profiles/r2_step_kernel_k{1,4,16}_ncu_full.txt); bench.py uses them for '
                     'roofline.traffic and roofline_k*'}
    for k in (1, 4, 16):
        path = os.path.join(src, 'r2_step_kernel_k%d_ncu_full.txt' % k)
        if not os.path.exists(path):
            continue
        m = first_launch(path)
        cap = {'kernel': m.get('kernel'), 'envs': envs, 'k': k,
               'batch': 'desynchronised (997 device-side steps on the --random stream first)',
               'duration_ms': scaled(m, 'gpu__time_duration.sum'), 'inst_executed': m['smsp__inst_executed.sum'][0],
               'issue_active_pct': m['smsp__issue_active.avg.pct_of_peak_sustained_active'][0],
               'registers': m['launch__registers_per_thread'][0], 'sm_ghz': scaled(m, 'gpc__cycles_elapsed.avg.per_second')}
        if k == 1:
            facts['k1_traffic'] = {'envs': envs, 'dram_bytes_per_launch': scaled(m, 'dram__bytes_read.sum') + scaled(m, 'dram__bytes_write.sum'),
                                   'duration_us': cap['duration_ms'] * 1e3,
                                   'capture': 'profiles/r2_step_kernel_k1_ncu_full.txt (desynchronised batch, --random stream)'}
            facts['k1_capture'] = cap
        else:
            facts['k%d_warp_instr_per_32_env_substeps' % k] = cap['inst_executed'] / (envs / 32 * k)
            facts['k%d_capture' % k] = cap
    with open(dst, 'w') as f:
        json.dump(facts, f, indent=1)
        f.write('\n')
    print(json.dumps({key: facts[key] for key in facts if not key.endswith('capture') and key != 'note'}))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out'),
         sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'profiles', 'r2_profile_facts.json'))
