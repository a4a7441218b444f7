cd $GRAFT_REPO_ROOT
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v25.log 2>&1; echo "pytest rc=$?"; tail -c 300 gpurun_out/pytest_gpu_v25.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v25.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_v25.log
timeout 400 python bench.py > gpurun_out/bench_v25.json 2> gpurun_out/bench_v25.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_v25.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['value_eager'], d['parity']['rel_diff'], d['exact_cmp_n32k']['seconds'], d['roofline']['frac'])"
