timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r32.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_r32.log
timeout 300 python tools/pipe_experiment.py > gpurun_out/pipe_sweep_r32.txt 2>&1; cat gpurun_out/pipe_sweep_r32.txt
