mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops_tc.py tests/test_gpu_ops.py -q -x -m gpu 2>&1 | tail -n 6
python tools/profile_step.py 8 bf16 kl488 2>&1 | head -n 24
VT_AB_LIB=/root/repo/vidtok_b200/libvt_ab2.so python tools/profile_step.py 8 bf16 kl488 2>&1 | head -n 12
python tools/profile_step.py 8 exact kl488 2>&1 | head -n 12
timeout 900 python -m pytest tests/test_gpu_model.py -q -x -m gpu 2>&1 | tail -n 4
