cd $GRAFT_REPO_ROOT
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v26.log 2>&1; echo "pytest rc=$?"; tail -c 300 gpurun_out/pytest_gpu_v26.log
S=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_v26.json 2> gpurun_out/bench_2gpu_v26.err; echo "bench2 rc=$? t=$(( $(date +%s)-S ))"
python -c "
import json; d=json.loads(open('gpurun_out/bench_2gpu_v26.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['value_eager'])"
grep -i "warn\|error" gpurun_out/bench_2gpu_v26.err | head -5
timeout 200 python tools/exact_sweep.py gpurun_out/exact_sweep_v26.json 16384 > gpurun_out/exact_sweep_v26.log 2>&1; echo "sweep rc=$?"
