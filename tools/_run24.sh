timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r24.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r24.log
timeout 300 python tools/pipe_experiment.py > gpurun_out/pipe_sweep_r24.txt 2>&1; cat gpurun_out/pipe_sweep_r24.txt
