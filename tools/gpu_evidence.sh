# N3 evidence pass, bounded: a few launches per kernel type (ncu --set full costs ~5 s per launch), each capture summarised on the
# box and copied to gpurun_out/profiles_r2 immediately; reports other than the conv_tc one are deleted (64 MiB return limit).
mkdir -p gpurun_out/profiles_r2
NCU="ncu --clock-control none"
cap() {  # tag regex skip count  B prec config  seconds
  timeout $8 $NCU --set full -k "regex:$2" -s $3 -c $4 -o gpurun_out/cap_$1 -f python tools/ncu_target.py $5 1 $6 $7 > gpurun_out/ncu_cap_$1.log 2>&1
  python tools/ncu_kernels.py gpurun_out/cap_$1.ncu-rep r2_$1 "\`ncu --set full --clock-control none -k regex:$2 -s $3 -c $4\` over \`python tools/ncu_target.py $5 1 $6 $7\` ($5 clip(s), precision $6, config $7)" > /dev/null
  rm -f gpurun_out/cap_$1.ncu-rep
  cp profiles/ncu_*_r2_$1.md profiles/ncu_index_r2_$1.md gpurun_out/profiles_r2/ 2>/dev/null
  echo "cap $1 done: $(ls gpurun_out/profiles_r2 | wc -l) files"
}
# launch list of the default bench command (metrics-only pass: cheap)
timeout 500 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list: $(wc -l < gpurun_out/launches_r2.csv) lines"
# the dominant kernel at the bench batch, with source (bench.py's roofline.traffic comes from this)
timeout 400 $NCU --set full --import-source on -k regex:conv_tc_kernel -s 100 -c 4 -o gpurun_out/prof_conv_tc_r2 -f python tools/ncu_target.py 8 1 > gpurun_out/ncu_conv_tc_r2.log 2>&1
ls -la gpurun_out/prof_conv_tc_r2.ncu-rep
cap tblock 'tblock2_tc_kernel' 1 2 8 bf16 kl488 300
cap misc_bf16 'conv_stem|conv_simt|tap_planes|softmax|transpose_bf16|ncdhw' 0 8 2 bf16 kl488 300
cap ln_bf16 'layernorm' 3 3 2 bf16 kl488 200
cap convs_bf16 'conv_tc_kernel' 10 10 2 bf16 kl488 300
cap exact 'conv_tc_kernel|layernorm_split|conv_stem|transpose_split|split_to_f32|f32_to_split' 12 12 1 exact kl488 300
cap fsq 'fsq|kl_|conv_simt|groupnorm' 0 8 1 mixed fsq488 300
cap v11 'time_interp|cache_update|copy_frames|upsample_nearest|stem_cache' 0 8 1 bf16 v11long 300
cap pack 'pack_w|absmax|fill_identity' 0 6 1 bf16 kl488 200
ls gpurun_out/profiles_r2 | wc -l; du -sh gpurun_out
