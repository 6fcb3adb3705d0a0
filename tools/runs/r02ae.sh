./tools/microbench/dmma_issue | tee gpurun_out/r02ae_dmma_issue.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scatter_slab' --launch-skip 1 -c 1 -o /tmp/r02ad_slab python tools/gpu_probe.py config2:DGZ > gpurun_out/r02ad_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02ad_slab.ncu-rep > gpurun_out/r02ad_ncu_full_config2_DGZ_scatter_slab_summary.txt 2>&1
ncu -i /tmp/r02ad_slab.ncu-rep --page source --csv > gpurun_out/r02ad_slab_source.csv 2>/dev/null
ls -la gpurun_out/r02ad* /tmp/r02ad_slab.ncu-rep
