mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops_tc.py -q -m gpu -s > gpurun_out/r2_t_ops_tc.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s > gpurun_out/r2_t_model.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_full.py -q -m gpu -s > gpurun_out/r2_t_full.log 2>&1
grep -E "passed|failed" gpurun_out/r2_t_ops_tc.log gpurun_out/r2_t_model.log gpurun_out/r2_t_full.log
grep -E "max err" gpurun_out/r2_t_ops_tc.log | head -20
grep -E "exact:" gpurun_out/r2_t_model.log | head -20
grep -E "^\[config|^\.\[config|FAILED|Error" gpurun_out/r2_t_full.log | head -30
VT_TBLOCK=2 python tools/profile_step.py 8 bf16 kl488 2>&1 | grep -E "total|tblock"
python tools/profile_step.py 8 exact kl488 > gpurun_out/r2_profile_step_exact.txt 2>&1; head -12 gpurun_out/r2_profile_step_exact.txt
