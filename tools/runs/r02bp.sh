timeout 400 python -X faulthandler -m pytest tests -m gpu -x -q -k "moments_tensor_core_shapes" > gpurun_out/r02bp_pytest.log 2>&1; tail -12 gpurun_out/r02bp_pytest.log
