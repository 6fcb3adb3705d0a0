timeout 600 ncu --set full --clock-control none --import-source on -k regex:'moments_mma' --launch-skip 2 -c 2 -o /tmp/r02al_mom python tools/gpu_probe.py config2:DGZ > gpurun_out/r02al_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02al_mom.ncu-rep > gpurun_out/r02al_ncu_full_config2_DGZ_moments_summary.txt 2>&1
ncu -i /tmp/r02al_mom.ncu-rep --page source --csv > gpurun_out/r02al_mom_source.csv 2>/dev/null
ls -la gpurun_out/r02al*
