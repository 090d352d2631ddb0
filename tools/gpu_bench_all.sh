# All four bench configurations + precision variants + the reference arm, one JSON line each -> gpurun_out/bench_r2_*.json
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_r2_$name.json 2> gpurun_out/bench_r2_$name.err; tail -c 600 gpurun_out/bench_r2_$name.json; echo; }
run kl488 --steps 20 --warmup 5
run kl488_exact --config kl488 --precision exact --steps 5 --warmup 3 --no-cpu-baseline
run fsq488 --config fsq488 --steps 10 --warmup 3 --no-cpu-baseline
run v11long --config v11long --steps 5 --warmup 3 --no-cpu-baseline
run kl41616 --config kl41616 --steps 10 --warmup 3 --no-cpu-baseline
run reference --impl reference --steps 2 --warmup 1
