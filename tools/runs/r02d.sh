TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sweep_pencil' --launch-skip 11 -c 11 -o gpurun_out/${TAG}_full_config2_ZGD_sweep_pencil python tools/gpu_probe.py config2:ZGD > gpurun_out/${TAG}_full_zgd.log 2>&1
tail -2 gpurun_out/${TAG}_full_zgd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sweep_pencil' --launch-skip 15 -c 3 -o gpurun_out/${TAG}_full_config2_GZD_sweep_pencil python tools/gpu_probe.py config2:GZD > gpurun_out/${TAG}_full_gzd.log 2>&1
tail -2 gpurun_out/${TAG}_full_gzd.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config2_DGZ.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_config2_DGZ.json; tail -3 gpurun_out/${TAG}_bench.err
