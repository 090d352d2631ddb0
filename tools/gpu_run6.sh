mkdir -p gpurun_out
for r in 0 1; do
  echo "== relaxed=$r"; VT_TB2_RELAXED=$r python tools/run_tblock.py 8 4 1; VT_TB2_RELAXED=$r python tools/run_tblock.py 8 3 0
  VT_TB2_RELAXED=$r timeout 600 python -m pytest tests/test_gpu_ops_tc.py -q -x -m gpu -k tblock 2>&1 | tail -n 3
done
timeout 900 python -m pytest tests/test_gpu_ops_tc.py tests/test_gpu_model.py -q -x -m gpu 2>&1 | tail -n 8
python tools/profile_step.py 8 bf16 kl488 2>&1 | head -n 14
python tools/profile_step.py 8 exact kl488 2>&1 | head -n 24
