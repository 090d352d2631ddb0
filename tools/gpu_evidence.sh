# N3 evidence pass (run under gpurun): launch list of the default bench, ncu --set full of every kernel of the four bench
# configurations (summarised on the box: the reports are too big to bring back), and the 4-launch conv_tc capture bench.py cites.
set -x
mkdir -p gpurun_out/profiles_r2
NCU="ncu --clock-control none"
# 1. launch list of the default bench command
timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# 2. every kernel, ncu --set full, per configuration
cap() {  # tag B prec config count
  timeout 1200 $NCU --set full -c $5 -o gpurun_out/all_$1 -f python tools/ncu_target.py $2 1 $3 $4 > gpurun_out/ncu_all_$1.log 2>&1
  python tools/ncu_kernels.py gpurun_out/all_$1.ncu-rep r2_$1 "\`ncu --set full --clock-control none -c $5\` over \`python tools/ncu_target.py $2 1 $3 $4\` (one forward, $2 clip(s), precision $3, config $4)"
  rm -f gpurun_out/all_$1.ncu-rep
}
cap bf16 2 bf16 kl488 400
cap exact 1 exact kl488 400
cap fsq 1 mixed fsq488 400
cap v11 1 bf16 v11long 500
cp profiles/ncu_*_r2_*.md profiles/ncu_index_r2_*.md gpurun_out/profiles_r2/ 2>/dev/null
# 3. the dominant kernel at the bench batch (bench.py reads profiles/ncu_conv_tc_r2.json built from this)
timeout 900 $NCU --set full --import-source on -k regex:conv_tc_kernel -s 100 -c 4 -o gpurun_out/prof_conv_tc_r2 -f python tools/ncu_target.py 8 1 > gpurun_out/ncu_conv_tc_r2.log 2>&1
ls -la gpurun_out gpurun_out/profiles_r2 | tail -n 40
