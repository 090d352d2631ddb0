mkdir -p gpurun_out
python tools/run_tblock.py 8 3 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tblock2 -s 1 -c 1 -o gpurun_out/tblock2_r2 -f python tools/run_tblock.py 8 2 1 > gpurun_out/ncu_tblock2.log 2>&1
tail -n 3 gpurun_out/ncu_tblock2.log
ls -la gpurun_out/tblock2_r2.ncu-rep
