timeout 200 python -X faulthandler -m pytest tests -m gpu -x -q -k "moments_tensor_core_shapes or lplustimes or ltimes" > gpurun_out/r02bq_pytest.log 2>&1; tail -15 gpurun_out/r02bq_pytest.log
