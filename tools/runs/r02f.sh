TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python tools/gpu_probe.py config2:GZD config2:ZGD > gpurun_out/${TAG}_probe.log 2>&1
grep -E "config|Sweep" gpurun_out/${TAG}_probe.log
timeout 300 python tools/gpu_probe.py config4:GZD >> gpurun_out/${TAG}_probe.log 2>&1
tail -8 gpurun_out/${TAG}_probe.log
