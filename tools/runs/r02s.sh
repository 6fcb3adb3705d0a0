TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/gpu_probe.py config2:GZD config2:ZGD config3:DGZ config4:DGZ 2>&1 | tee gpurun_out/${TAG}_probe.log | grep -vE "dfma"
