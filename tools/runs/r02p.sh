TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -c 3000 gpurun_out/${TAG}_pytest_gpu.log
