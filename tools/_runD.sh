timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r34.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r34.log
( timeout 100 python tools/profile_kernels.py B=256 1; timeout 100 python tools/profile_kernels.py 500 ) > gpurun_out/prof_r34.txt 2>&1; cat gpurun_out/prof_r34.txt
