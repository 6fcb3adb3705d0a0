for e in 0 1 2 3 4 6 7; do
  echo "KB200_SLAB_EXP=$e"
  KB200_SLAB_EXP=$e timeout 300 python tools/gpu_probe.py config2:DGZ config3:DGZ 2>&1 | grep -E "scatt"
done | tee gpurun_out/r02ad_slab_experiments.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scatter_slab' --launch-skip 3 -c 1 -o /tmp/r02ad_slab python tools/gpu_probe.py config2:DGZ > gpurun_out/r02ad_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02ad_slab.ncu-rep > gpurun_out/r02ad_ncu_full_config2_DGZ_scatter_slab_summary.txt 2>&1
ncu -i /tmp/r02ad_slab.ncu-rep --page source --csv > gpurun_out/r02ad_slab_source.csv 2>/dev/null
ls -la gpurun_out/r02ad* /tmp/r02ad_slab.ncu-rep
