cd $GRAFT_REPO_ROOT
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v23.log 2>&1; echo "pytest rc=$?"; tail -c 300 gpurun_out/pytest_gpu_v23.log
timeout 300 python tools/exact_sweep.py gpurun_out/exact_sweep_v23.json > gpurun_out/exact_sweep_v23.log 2>&1; echo "sweep rc=$?"
timeout 120 python tools/one_potrf.py f32 8192 > gpurun_out/one_potrf_v23.log 2>&1; echo "one_potrf rc=$?"; cat gpurun_out/one_potrf_v23.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:potrf_step_kernel -s 20 -c 2 -f -o gpurun_out/potrf_step_v23 python tools/one_potrf.py f32 8192 > gpurun_out/ncu_potrf_step_v23.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_potrf_step_v23.log; ls -la gpurun_out/potrf_step_v23.ncu-rep
