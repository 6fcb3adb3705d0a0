timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lplustimes_slab|ltimes_slab' --launch-skip 2 -c 2 -o /tmp/r02bf python tools/gpu_probe.py config2:DGZ > gpurun_out/r02bf_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02bf.ncu-rep > gpurun_out/r02bf_ncu_full_config2_DGZ_moments_slab_summary.txt 2>&1
ncu -i /tmp/r02bf.ncu-rep --page source --csv > gpurun_out/r02bf_source.csv 2>/dev/null
