set -x
cd $GRAFT_REPO_ROOT
S=$(date +%s)
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v20.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" | tee -a gpurun_out/pytest_gpu_v20.log
S=$(date +%s)
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v20.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s)-S ))" | tee -a gpurun_out/smoke_v20.log
S=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_v20.json 2> gpurun_out/bench_v20.err; echo "bench rc=$? t=$(( $(date +%s)-S ))"
S=$(date +%s)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_v20.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-exact > gpurun_out/ncu_v20.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s)-S ))"
tail -c 600 gpurun_out/pytest_gpu_v20.log
