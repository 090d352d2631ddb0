mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops_tc.py -q -m gpu -s 2>&1 > gpurun_out/r2_t_ops_tc.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu 2>&1 | tail -n 40 > gpurun_out/r2_t_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s 2>&1 | tail -n 120 > gpurun_out/r2_t_model.log
timeout 1500 python -m pytest tests/test_gpu_full.py -q -m gpu -s -x 2>&1 | tail -n 120 > gpurun_out/r2_t_full.log
grep -E "passed|failed" gpurun_out/r2_t_ops_tc.log gpurun_out/r2_t_ops.log gpurun_out/r2_t_model.log gpurun_out/r2_t_full.log
