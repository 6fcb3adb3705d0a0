TAG=${1:-r02o}
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -30 gpurun_out/${TAG}_pytest_gpu.log
for x in 0 1 2 3 4 7; do
  echo "== PENCIL_EXP=$x"
  KB200_PENCIL_EXP=$x timeout 300 python tools/gpu_probe.py config2:GZD config2:ZGD 2>&1 | grep -E "config2|Sweep"
done
