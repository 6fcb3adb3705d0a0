timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded" > gpurun_out/r02ac_pytest_scat.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02ac_pytest_scat.log
tail -15 gpurun_out/r02ac_pytest_scat.log
timeout 300 python tools/gpu_probe.py config2:DGZ config3:DGZ config4:DGZ config2:ZGD 2>&1 | grep -E "config|scatt|Sweep" | tee gpurun_out/r02ac_probe.log
