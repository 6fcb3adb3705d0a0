TAG=${1:-r02g}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "sweep or solve or golden or invariance or population" 2>&1 | tail -15; echo "pytest rc=$?") > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python tools/gpu_probe.py config2:GZD config2:ZGD > gpurun_out/${TAG}_probe.log 2>&1
grep -E "config|Sweep" gpurun_out/${TAG}_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sweep_pencil' --launch-skip 15 -c 2 -o gpurun_out/${TAG}_full_config2_ZGD_sweep_pencil python tools/gpu_probe.py config2:ZGD > gpurun_out/${TAG}_full_zgd.log 2>&1
tail -2 gpurun_out/${TAG}_full_zgd.log
