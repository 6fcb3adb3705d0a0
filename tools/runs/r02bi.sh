timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded or ltimes_abi" 2>&1 | tail -3
for e in 64 16; do echo KB200_ZGD_ZT=$e; for c in config2:ZGD config4:ZGD config3:ZGD; do
  KB200_ZGD_ZT=$e timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|scatt|fill_gbs"
done; done | tee gpurun_out/r02bi_probe.log
