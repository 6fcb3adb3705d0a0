timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "lplustimes or moments or ltimes" 2>&1 | tail -2
for c in config2:DGZ config3:DGZ config4:DGZ; do
  timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|Times"
done | tee gpurun_out/r02am_probe.log
