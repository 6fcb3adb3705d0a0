set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sweep_zline|moments_mma|scatter_mma|population_vec' --launch-skip 9 -c 9 -o gpurun_out/full_config2_DGZ python tools/gpu_probe.py config2:DGZ > gpurun_out/full.log 2>&1
tail -3 gpurun_out/full.log
