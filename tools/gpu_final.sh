mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 600 python bench.py --config fsq488 --steps 5 --warmup 3 > gpurun_out/bench_r2_fsq488_parity.json 2> gpurun_out/bench_r2_fsq488_parity.err; tail -c 900 gpurun_out/bench_r2_fsq488_parity.json
