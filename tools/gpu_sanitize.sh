# compute-sanitizer memcheck over a few small cases of the new kernels (bounded: the sanitizer slows kernels 10-100x)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops_tc.py -q -x -m gpu \
  -k "(tiny_kl_v10-exact or tiny_fsq_v11_tiled-exact or tiny_kl_nc-exact or host_staging) or (temporal_resblock and w16) or (fsq_epilogue) or (kl_epilogue and 4-)" \
  > gpurun_out/sanitize_r2.log 2>&1
echo "rc=$?"
tail -n 15 gpurun_out/sanitize_r2.log
