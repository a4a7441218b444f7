cd $GRAFT_REPO_ROOT
S=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_v20.json 2> gpurun_out/bench_2gpu_v20.err; echo "bench2 rc=$? t=$(( $(date +%s)-S ))" | tee gpurun_out/bench_2gpu_v20.rc
S=$(date +%s)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu_v20.json 2> gpurun_out/bench_ref_2gpu_v20.err; echo "ref2 rc=$? t=$(( $(date +%s)-S ))" | tee -a gpurun_out/bench_2gpu_v20.rc
tail -c 300 gpurun_out/bench_2gpu_v20.json
