timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded" 2>&1 | tail -2
for c in config2:ZGD config4:ZGD config3:ZGD config3:GZD config4:GZD; do
  timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|scatt"
done | tee gpurun_out/r02ao_probe.log
