timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "ltimes or lplustimes or moments" > gpurun_out/r02bu_pytest.log 2>&1; tail -3 gpurun_out/r02bu_pytest.log
