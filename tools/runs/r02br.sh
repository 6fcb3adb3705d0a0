timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r02br_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02br_pytest_gpu.log
tail -3 gpurun_out/r02br_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none -k regex:'scatter_slab|moments_mma' --launch-skip 3 -c 3 -o /tmp/r02br_c3 python tools/gpu_probe.py config3:DGZ > gpurun_out/r02br_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02br_c3.ncu-rep > gpurun_out/r02br_ncu_full_config3_DGZ_summary.txt 2>&1
timeout 120 python tools/gpu_probe.py config3:DGZ config2:DGZ 2>&1 | tee gpurun_out/r02br_probe.log | grep -E "config"
