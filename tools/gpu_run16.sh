mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b16_kl488.json 2> gpurun_out/b16_kl488.err; tail -c 1200 gpurun_out/b16_kl488.json; tail -n 3 gpurun_out/b16_kl488.err
timeout 600 python bench.py --config v11long --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b16_v11.json 2> gpurun_out/b16_v11.err; tail -c 900 gpurun_out/b16_v11.json; tail -n 3 gpurun_out/b16_v11.err
