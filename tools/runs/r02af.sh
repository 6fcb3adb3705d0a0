timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded" > gpurun_out/r02af_pytest_scat.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02af_pytest_scat.log
tail -4 gpurun_out/r02af_pytest_scat.log
for e in 0 1 2 6; do
  echo "KB200_SLAB_EXP=$e"
  KB200_SLAB_EXP=$e timeout 300 python tools/gpu_probe.py config2:DGZ config3:DGZ config4:DGZ 2>&1 | grep -E "scatt"
done | tee gpurun_out/r02af_slab_experiments.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scatter_slab' --launch-skip 1 -c 1 -o /tmp/r02af_slab python tools/gpu_probe.py config2:DGZ > gpurun_out/r02af_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02af_slab.ncu-rep > gpurun_out/r02af_ncu_full_config2_DGZ_scatter_slab_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scatter_slab' --launch-skip 1 -c 1 -o /tmp/r02af_slab3 python tools/gpu_probe.py config3:DGZ > gpurun_out/r02af_ncu3.log 2>&1
python tools/ncu_summary.py /tmp/r02af_slab3.ncu-rep > gpurun_out/r02af_ncu_full_config3_DGZ_scatter_slab_summary.txt 2>&1
