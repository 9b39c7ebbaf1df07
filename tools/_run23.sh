timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r23.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r23.log
( timeout 120 python tools/profile_kernels.py 1
  timeout 120 python tools/profile_kernels.py B=256 1 ) > gpurun_out/prof_r23.txt 2>&1
cat gpurun_out/prof_r23.txt
timeout 300 python tools/pipe_experiment.py > gpurun_out/pipe_sweep_r23.txt 2>&1; cat gpurun_out/pipe_sweep_r23.txt
