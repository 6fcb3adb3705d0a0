timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded" 2>&1 | tail -2
for e in 1 0; do
echo KB200_ZGD_TRANSPOSE=$e
for c in config2:ZGD config4:ZGD config3:ZGD; do
  KB200_ZGD_TRANSPOSE=$e timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|scatt"
done; done | tee gpurun_out/r02an_probe.log
