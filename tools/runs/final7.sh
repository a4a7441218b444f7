cd $GRAFT_REPO_ROOT
S=$(date +%s)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_v27.json 2> gpurun_out/bench_8gpu_v27.err; echo "bench8 rc=$? t=$(( $(date +%s)-S ))"
python -c "
import json; d=json.loads(open('gpurun_out/bench_8gpu_v27.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['value_eager'])"
grep -i "warn\|error" gpurun_out/bench_8gpu_v27.err | head -5
