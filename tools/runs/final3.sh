cd $GRAFT_REPO_ROOT
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v22.log 2>&1; echo "pytest rc=$?"; tail -c 400 gpurun_out/pytest_gpu_v22.log
timeout 400 python bench.py > gpurun_out/bench_v22.json 2> gpurun_out/bench_v22.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_v22.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['value_eager'], d['parity']['rel_diff'])"
timeout 300 python tools/exact_sweep.py gpurun_out/exact_sweep_v22.json > gpurun_out/exact_sweep_v22.log 2>&1; echo "sweep rc=$?"; tail -c 1500 gpurun_out/exact_sweep_v22.log
