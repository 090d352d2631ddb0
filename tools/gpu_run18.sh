mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_full.py -q -x -m gpu -k "v11 or tiled or config4 or video or chunk" 2>&1 | tail -n 4
python tools/profile_step.py 1 bf16 v11long 2>&1 | grep -E "total|cache_update|copy_frames|layernorm|time_interp"
timeout 600 python bench.py --config v11long --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_v11long_b.json 2> gpurun_out/b18_v11.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_v11long_b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['profiled'], d['roofline']['kernels_ms'])
PY
