timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r25.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r25.log
timeout 100 python tools/profile_kernels.py 1 500 > gpurun_out/prof_r25.txt 2>&1; cat gpurun_out/prof_r25.txt
timeout 300 python tools/pipe_experiment.py > gpurun_out/pipe_sweep_r25.txt 2>&1; cat gpurun_out/pipe_sweep_r25.txt
