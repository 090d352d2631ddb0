# the round-end checks the driver runs, in one call: GPU test suite, smoke(), a short default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 700 gpurun_out/bench_final.json; tail -n 2 gpurun_out/bench_final.err
