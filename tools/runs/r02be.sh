for e in 1 0; do echo KB200_MS_ADJ=$e; for c in config2:DGZ config4:DGZ; do
  KB200_MS_ADJ=$e timeout 60 python tools/gpu_probe.py $c 2>&1 | grep -E "config|Times"
done; done | tee gpurun_out/r02be_probe.log
