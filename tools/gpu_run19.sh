mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -x -m gpu -k "v11" 2>&1 | tail -n 4
timeout 600 python bench.py --config v11long --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_v11long_c.json 2> gpurun_out/b19_v11.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_v11long_c.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['profiled'], d['roofline']['kernels_ms'], d['psnr'])
PY
