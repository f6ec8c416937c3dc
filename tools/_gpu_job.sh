mkdir -p gpurun_out
COPTER_HYP_EXAMPLES=600 COPTER_HYP_RANDOM=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k test_random_shapes_vs_oracle 2>&1 | tail -40 > gpurun_out/hyp_many.log; tail -5 gpurun_out/hyp_many.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k test_random_shapes_vs_oracle 2>&1 | tail -3
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2_tracker_flips.jsonl')]
rows=[r for r in rows if not r['fp64']]
rows.sort(key=lambda r:-r['flips'])
print(len(rows),'fp32 examples; top flips:',[(r['flips'],r['episodes']) for r in rows[:12]])
PY
