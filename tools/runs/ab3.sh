# GPU tests of the runner launcher, then A/B of the side-stream K_ZZ factorisation under graph replay
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_runner_gpu.py -x -q > gpurun_out/pytest_runner_v21.log 2>&1; echo "runner tests rc=$?"; tail -c 1500 gpurun_out/pytest_runner_v21.log
: > gpurun_out/ab_side_v21.txt
for rep in 1 2; do
for cfg in "X=0" "DCGP_SIDE_STREAM=1" "DCGP_SIDE_STREAM=1 DCGP_TC_RESERVE=2" "DCGP_TC_RESERVE=2" "DCGP_SIDE_STREAM=1 DCGP_TC_RESERVE=3"; do
  env $cfg timeout 300 python bench.py --no-cpu --no-exact 2>/tmp/ab.err > /tmp/ab.json
  python - "$cfg" <<'PY' | tee -a gpurun_out/ab_side_v21.txt
import json, sys
try:
    d = json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
    print(sys.argv[1], '|', round(d['value'], 2), 'steps/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 2), 'eager', round(d['value_eager'], 2), 'loss', d['final_loss'], d['config']['step_mode'][:10])
except Exception as e:
    print(sys.argv[1], '| FAILED', e, open('/tmp/ab.err').read()[-500:])
PY
  grep -i "warn\|fail" /tmp/ab.err | head -3 | tee -a gpurun_out/ab_side_v21.txt
done
done
