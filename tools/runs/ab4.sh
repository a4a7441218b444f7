cd $GRAFT_REPO_ROOT
DCGP_SIDE_WHITEN=1 timeout 300 python -m pytest tests/test_models_gpu.py -x -q -k "graph or side or trainer" > gpurun_out/pytest_side_whiten_v24.log 2>&1; echo "tests rc=$?"; tail -c 600 gpurun_out/pytest_side_whiten_v24.log
: > gpurun_out/ab_side_whiten_v24.txt
for rep in 1 2 3; do
for cfg in "DCGP_SIDE_WHITEN=0" "DCGP_SIDE_WHITEN=1"; do
  env $cfg timeout 300 python bench.py --no-cpu --no-exact 2>/tmp/ab.err > /tmp/ab.json
  python - "$cfg" <<'PY' | tee -a gpurun_out/ab_side_whiten_v24.txt
import json, sys
try:
    d = json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
    print(sys.argv[1], '|', round(d['value'], 2), 'steps/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 2), 'eager', round(d['value_eager'], 2), 'loss', d['final_loss'], d['config']['step_mode'][:10])
except Exception as e:
    print(sys.argv[1], '| FAILED', e, open('/tmp/ab.err').read()[-500:])
PY
  grep -i "warn\|fail" /tmp/ab.err | head -3 | tee -a gpurun_out/ab_side_whiten_v24.txt
done
done
