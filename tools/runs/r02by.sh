timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "scattering or folded" > gpurun_out/r02by_pytest.log 2>&1; tail -3 gpurun_out/r02by_pytest.log
timeout 60 python tools/gpu_probe.py config2:DGZ 2>&1 | grep -E "config|scatt"
