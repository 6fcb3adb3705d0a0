timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "lplustimes or moments or full_size" > gpurun_out/r02bo_pytest.log 2>&1; tail -3 gpurun_out/r02bo_pytest.log
for e in 1 0; do echo KB200_GEMM_SLAB=$e; for c in config3:DGZ; do
  KB200_GEMM_SLAB=$e timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|Times"
done; done | tee gpurun_out/r02bo_probe.log
