mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops_tc.py -q -x -m gpu 2>&1 | tail -n 6
for v in 2 0; do echo "== VT_TBLOCK=$v"; VT_TBLOCK=$v python tools/profile_step.py 8 bf16 kl488 2>&1 | grep -E "total|tblock|k311 s11 128->128" ; done
python tools/profile_step.py 8 exact kl488 2>&1 | head -n 30
timeout 900 python -m pytest tests/test_gpu_model.py -q -x -m gpu 2>&1 | tail -n 4
