mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops_tc.py -q -m gpu 2>&1 | tail -n 30 > gpurun_out/r2_t_ops_tc.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "video or mixed or autocast" -s 2>&1 | tail -n 60 > gpurun_out/r2_t_model2.log
timeout 1500 python -m pytest tests/test_gpu_full.py -q -m gpu -s -k "config3 or config4 or config5" 2>&1 | tail -n 120 > gpurun_out/r2_t_full.log
grep -E "passed|failed" gpurun_out/r2_t_ops_tc.log gpurun_out/r2_t_model2.log gpurun_out/r2_t_full.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_kl488.json 2> gpurun_out/r2_bench_kl488.err
VT_TBLOCK=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_kl488_notblock.json 2> gpurun_out/r2_bench_kl488_notblock.err
python bench.py --steps 4 --warmup 3 --precision exact --no-cpu-baseline > gpurun_out/r2_bench_kl488_exact.json 2> gpurun_out/r2_bench_kl488_exact.err
python bench.py --steps 6 --warmup 3 --config fsq488 > gpurun_out/r2_bench_fsq488.json 2> gpurun_out/r2_bench_fsq488.err
python bench.py --steps 6 --warmup 3 --config v11long > gpurun_out/r2_bench_v11long.json 2> gpurun_out/r2_bench_v11long.err
python bench.py --steps 6 --warmup 3 --config kl41616 > gpurun_out/r2_bench_kl41616.json 2> gpurun_out/r2_bench_kl41616.err
python tools/profile_step.py 8 bf16 kl488 > gpurun_out/r2_profile_step_bf16.txt 2>&1
python tools/profile_step.py 8 exact kl488 > gpurun_out/r2_profile_step_exact.txt 2>&1
python tools/profile_step.py 1 bf16 v11long > gpurun_out/r2_profile_step_v11long.txt 2>&1
for f in gpurun_out/r2_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(" value %.1f  e2e %.1f  ms/step %.2f  %s frac %.3f  sumk %.1f  launches %s  clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel"), r.get("frac",0), r.get("sum_kernel_ms",0), d.get("gpu_launches"), d.get("clocks")))
    print(" psnr", d.get("psnr"))
except Exception as e:
    print(" parse failed", e)
PY
done
tail -n 3 gpurun_out/r2_bench_*.err
