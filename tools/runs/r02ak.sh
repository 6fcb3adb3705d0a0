timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded" 2>&1 | tail -2
for c in config2:DGZ config3:DGZ config4:DGZ config2:ZGD config1:DGZ config5:DGZ; do
  timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|scatt"
done | tee gpurun_out/r02ak_probe.log
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r02ak_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02ak_pytest_gpu.log
tail -3 gpurun_out/r02ak_pytest_gpu.log
