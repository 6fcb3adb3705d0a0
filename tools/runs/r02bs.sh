timeout 400 python -X faulthandler -m pytest tests -m gpu -x -q -k "moments_tensor_core_shapes or ltimes or full_size" > gpurun_out/r02bs_pytest.log 2>&1; tail -4 gpurun_out/r02bs_pytest.log
for e in 1 0; do echo KB200_LTIMES_SLAB=$e; KB200_LTIMES_SLAB=$e timeout 60 python tools/gpu_probe.py config3:DGZ 2>&1 | grep -E "config|LTimes"; done | tee gpurun_out/r02bs_probe.log
