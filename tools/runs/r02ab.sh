timeout 300 python tools/gpu_probe.py config2:DGZ config3:DGZ 2>&1 | grep -E "config|scatt|Sweep"
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r02ab_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02ab_pytest_gpu.log
tail -3 gpurun_out/r02ab_pytest_gpu.log
