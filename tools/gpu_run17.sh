mkdir -p gpurun_out
python tools/profile_step.py 1 bf16 v11long 2>&1 | head -n 28
timeout 600 python bench.py --config v11long --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b17_v11.json 2> gpurun_out/b17_v11.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b17_v11.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['profiled'], d['roofline']['kernels_ms'])
PY
bash tools/gpu_sanitize.sh
