timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded or ltimes_abi" > gpurun_out/r02bj_pytest.log 2>&1; tail -3 gpurun_out/r02bj_pytest.log
for c in config2:ZGD config4:ZGD; do
  timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|scatt"
done | tee gpurun_out/r02bj_probe.log
