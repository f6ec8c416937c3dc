"""Static checks on the SASS of the built library (no GPU needed: cuobjdump reads the cubin).

The step kernel is memory-bound, so each thread must have ALL of its input loads (state planes,
meta word, action row) in flight before its first stall.  ptxas has, after unrelated edits, placed
the clip of the action row between the action loads and the state loads (fp64 Hover3D: 0.221
instead of 0.190 ms per launch on B200); `load_raw` now ties the action to the other inputs, and
this test pins the resulting schedule: the input loads of every `copter_step_kernel` instantiation
sit within a few instructions of each other."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACTION = {0: 4, 1: 2, 2: 1, 3: 4, 4: 2, 5: 1, 6: 4}           # variant id -> action size (include/copter_b200.h)


@pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='cuobjdump not on PATH')
def test_step_kernel_issues_all_input_loads_together():
    from gym_copter_b200 import build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    seen = 0
    for fn in re.split(r'\n\s*Function : ', sass)[1:]:
        m = re.search(r'copter_step_kernelI([fd])Li(\d)ELb([01])', fn.split('\n', 1)[0])
        if not m:
            continue
        ops = [re.sub(r'^\s+/\*[0-9a-f]+\*/\s+', '', line).split('/*')[0].strip()
               for line in fn.split('\n') if re.match(r'\s+/\*[0-9a-f]{4}\*/', line)]
        word = 4 if m.group(1) == 'f' else 8
        n_inputs = 12 * word // 16 + 1 + max(1, ACTION[int(m.group(2))] * word // 16)
        loads = [i for i, op in enumerate(ops) if op.startswith('LDG')][:n_inputs]       # unpredicated: the inputs come first
        assert len(loads) == n_inputs, m.group(0)
        window = loads[-1] - loads[0]
        between = ops[loads[0]:loads[-1]]
        assert window <= 32, (m.group(0), window)
        # nothing in between may wait on a loaded value: no floating-point compare / min-max (the clip)
        assert not [op for op in between if re.match(r'(@!?U?P\d+\s+)?(DSETP|FSETP|FMNMX|DMNMX)', op)], m.group(0)
        seen += 1
    assert seen == 28          # 7 variants x 2 precisions x with/without statistics
