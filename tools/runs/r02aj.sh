for w in 2 4; do
KB200_SLAB_GROUPS=$w timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded" > gpurun_out/r02aj_pytest_scat_$w.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02aj_pytest_scat_$w.log
tail -2 gpurun_out/r02aj_pytest_scat_$w.log
done
for w in 2 4 1; do
  echo "KB200_SLAB_GROUPS=$w"
  for c in config2:DGZ config3:DGZ config4:DGZ; do
    KB200_SLAB_GROUPS=$w timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "scatt"
  done
done | tee gpurun_out/r02aj_slab_groups.log
