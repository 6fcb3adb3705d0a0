timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "ltimes or lplustimes or moments or golden or full_size" 2>&1 | tail -3
for e in 1 0; do echo KB200_LPLUSTIMES_SLAB=$e; for c in config2:DGZ config4:DGZ config5:DGZ config2:GDZ; do
  KB200_LPLUSTIMES_SLAB=$e timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|Times"
done; done | tee gpurun_out/r02bd_probe.log
