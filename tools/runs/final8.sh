cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v28.log 2>&1; echo "pytest rc=$?"; tail -c 200 gpurun_out/pytest_gpu_v28.log
DCGP_GRAM_RT=1 timeout 100 python tools/gram_bw.py > gpurun_out/gram_bw_v28.jsonl 2>gpurun_out/gram_bw_v28.err
timeout 100 python tools/gram_bw.py >> gpurun_out/gram_bw_v28.jsonl 2>>gpurun_out/gram_bw_v28.err
DCGP_GRAM_RT=32 timeout 100 python tools/gram_bw.py >> gpurun_out/gram_bw_v28.jsonl 2>>gpurun_out/gram_bw_v28.err
cat gpurun_out/gram_bw_v28.jsonl | cut -c1-200
timeout 200 python bench.py > gpurun_out/bench_v28.json 2> gpurun_out/bench_v28.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_v28.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['rel_diff'], d['exact_cmp_n32k']['seconds'], d['exact_cmp_n32k']['phases_ms'])"
