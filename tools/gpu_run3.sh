mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops_tc.py -q -m gpu -s 2>&1 | tail -n 150 > gpurun_out/r2_t_ops_tc.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu 2>&1 | tail -n 20 > gpurun_out/r2_t_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s 2>&1 | tail -n 150 > gpurun_out/r2_t_model.log
timeout 1500 python -m pytest tests/test_gpu_full.py -q -m gpu -s 2>&1 | tail -n 150 > gpurun_out/r2_t_full.log
grep -E "passed|failed" gpurun_out/r2_t_ops_tc.log gpurun_out/r2_t_ops.log gpurun_out/r2_t_model.log gpurun_out/r2_t_full.log
python bench.py --steps 4 --warmup 3 --precision exact --no-cpu-baseline > gpurun_out/r2_bench_kl488_exact.json 2> gpurun_out/r2_bench_kl488_exact.err
python bench.py --steps 6 --warmup 3 --config fsq488 > gpurun_out/r2_bench_fsq488.json 2> gpurun_out/r2_bench_fsq488.err
tail -c 1500 gpurun_out/r2_bench_fsq488.json
