set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r21.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_r21.log
( timeout 120 python tools/profile_kernels.py B=256 1
  L2H_LSTM_STAGED=0 timeout 120 python tools/profile_kernels.py B=256 1
  L2H_LSTM_PER=4 timeout 120 python tools/profile_kernels.py B=256 1
  L2H_LSTM_PER=1 timeout 120 python tools/profile_kernels.py B=256 1
  timeout 120 python tools/profile_kernels.py 500
  L2H_LSTM_STAGED=0 timeout 120 python tools/profile_kernels.py 500
  L2H_LSTM_PER=2 timeout 120 python tools/profile_kernels.py 500 ) > gpurun_out/prof_r21.txt 2>&1
cat gpurun_out/prof_r21.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rows_gemm_big|lstm_rec4|mid_kernel|attn_kernel" --launch-skip 12 -c 4 -f -o gpurun_out/prof_b256 python tools/prof_chain.py 1 256 3 > gpurun_out/ncu_b256.log 2>&1; tail -3 gpurun_out/ncu_b256.log
