cd $GRAFT_REPO_ROOT
timeout 100 python -m pytest tests/test_runner_gpu.py tests/test_models_gpu.py tests/test_compat_gpu.py -q -m gpu > gpurun_out/pytest_gpu_v29.log 2>&1; echo "pytest rc=$?"; tail -c 3500 gpurun_out/pytest_gpu_v29.log
