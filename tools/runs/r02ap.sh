for c in config3 config4 config5 config1; do for l in DGZ DZG GDZ GZD ZDG ZGD; do
  timeout 90 python tools/gpu_probe.py $c:$l 2>&1 | grep -vE "dfma_gflops"
done; done > gpurun_out/r02ap_probe_matrix.log 2>&1
grep -E "config" gpurun_out/r02ap_probe_matrix.log
