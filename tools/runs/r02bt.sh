timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r02bt_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02bt_pytest_gpu.log
tail -3 gpurun_out/r02bt_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02bt_bench_config2_DGZ.json 2> gpurun_out/r02bt_bench.err
cut -c1-300 gpurun_out/r02bt_bench_config2_DGZ.json
timeout 200 python tools/gpu_probe.py config2:DGZ config3:DGZ config4:DGZ config5:DGZ > gpurun_out/r02bt_probe.log 2>&1; grep config gpurun_out/r02bt_probe.log
