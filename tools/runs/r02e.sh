TAG=${1:-r02e}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25; echo "pytest rc=$?") > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -12 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python tools/gpu_probe.py config2:GZD config2:ZGD config1:ZGD config4:GZD > gpurun_out/${TAG}_probe.log 2>&1
cat gpurun_out/${TAG}_probe.log
