timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r02bz_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02bz_pytest_gpu.log
tail -3 gpurun_out/r02bz_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
