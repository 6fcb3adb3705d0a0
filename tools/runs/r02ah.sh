timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded" > gpurun_out/r02ah_pytest_scat.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02ah_pytest_scat.log
tail -4 gpurun_out/r02ah_pytest_scat.log
for w in 2 1; do
  echo "KB200_SLAB_GROUPS=$w"
  KB200_SLAB_GROUPS=$w timeout 300 python tools/gpu_probe.py config2:DGZ config3:DGZ config4:DGZ 2>&1 | grep -E "scatt"
done | tee gpurun_out/r02ah_slab_warps.log
KB200_SLAB_GROUPS=1 timeout 300 python -m pytest tests -m gpu -x -q -k "scattering or folded" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scatter_slab' --launch-skip 1 -c 1 -o /tmp/r02ah_slab python tools/gpu_probe.py config2:DGZ > gpurun_out/r02ah_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02ah_slab.ncu-rep > gpurun_out/r02ah_ncu_full_config2_DGZ_scatter_slab_summary.txt 2>&1
ncu -i /tmp/r02ah_slab.ncu-rep --page source --csv > gpurun_out/r02ah_slab_source.csv 2>/dev/null
