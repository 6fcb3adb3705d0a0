TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q -k "device_generator" > gpurun_out/${TAG}_pytest_gen.log 2>&1
tail -40 gpurun_out/${TAG}_pytest_gen.log
timeout 900 python tools/gpu_probe.py config2:DGZ > gpurun_out/${TAG}_probe.log 2>&1
grep -E "config|Sweep" gpurun_out/${TAG}_probe.log
KB200_IROW_SWIZZLE=0 timeout 900 python tools/gpu_probe.py config2:DGZ 2>&1 | grep -E "config|Sweep"
timeout 300 python -m pytest tests -m gpu -x -q -k "irow or zone_fastest or golden" 2>&1 | tail -3
