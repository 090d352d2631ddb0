mkdir -p gpurun_out
python tools/run_tblock.py 8 4 1; python tools/run_tblock.py 8 3 0
timeout 600 python -m pytest tests/test_gpu_ops_tc.py -q -x -m gpu -k temporal_resblock -s 2>&1 | tail -n 4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tblock2 -s 1 -c 1 -o gpurun_out/tblock2b_r2 -f python tools/run_tblock.py 8 2 1 > gpurun_out/ncu_tblock2b.log 2>&1
tail -n 2 gpurun_out/ncu_tblock2b.log
python tools/profile_step.py 8 bf16 kl488 2>&1 | head -n 8
