# N3 evidence pass, bounded: every capture copies its summaries to gpurun_out/ as soon as it is done.
# usage: bash tools/gpu_evidence2.sh [list] [bf16] [exact] [fsq] [v11] [conv]
mkdir -p gpurun_out/profiles_r2
NCU="ncu --clock-control none"
cap() {  # tag B prec config count seconds
  timeout $6 $NCU --set full -c $5 -o gpurun_out/all_$1 -f python tools/ncu_target.py $2 1 $3 $4 > gpurun_out/ncu_all_$1.log 2>&1
  python tools/ncu_kernels.py gpurun_out/all_$1.ncu-rep r2_$1 "\`ncu --set full --clock-control none -c $5\` over \`python tools/ncu_target.py $2 1 $3 $4\` (one forward, $2 clip(s), precision $3, config $4)"
  rm -f gpurun_out/all_$1.ncu-rep
  cp profiles/ncu_*_r2_$1.md profiles/ncu_index_r2_$1.md gpurun_out/profiles_r2/ 2>/dev/null
}
for what in "$@"; do
  case $what in
    list) timeout 600 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1 ;;
    bf16) cap bf16 1 bf16 kl488 400 500 ;;
    exact) cap exact 1 exact kl488 400 500 ;;
    fsq) cap fsq 1 mixed fsq488 400 500 ;;
    v11) cap v11 1 bf16 v11long 260 500 ;;
    conv) timeout 600 $NCU --set full --import-source on -k regex:conv_tc_kernel -s 100 -c 4 -o gpurun_out/prof_conv_tc_r2 -f python tools/ncu_target.py 8 1 > gpurun_out/ncu_conv_tc_r2.log 2>&1 ;;
  esac
done
ls gpurun_out/profiles_r2 | wc -l
